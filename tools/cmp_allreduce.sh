#!/bin/bash
# compare the fused peer-memory all-reduce with NCCL at N GPUs ($1)
n=$1
for mode in peer nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus $n --steps 30 --warmup 3 --allreduce $mode 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$mode', 'gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'max', max(d['step_ms']), 'argmax', d['step_ms'].index(max(d['step_ms'])), 'median', sorted(d['step_ms'])[len(d['step_ms'])//2], 'top3', sorted(d['step_ms'])[-3:], {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})"
done
