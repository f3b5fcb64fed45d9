#!/bin/bash
# 1/2/4/8-GPU strong-scaling run of the headline workload (what the driver does at round end)
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = "1" ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$n.json") if l.startswith("{")][-1])
    print("gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("gpus $n failed", e); print(open("gpurun_out/scale_$n.err").read()[-600:])
PY
done
