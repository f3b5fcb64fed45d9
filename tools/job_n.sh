mkdir -p gpurun_out; rm -f gpurun_out/multirank_parity.jsonl
nvidia-smi -L
python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -5
cp gpurun_out/multirank_parity.jsonl gpurun_out/r02_multirank_parity_2gpu.jsonl; cat gpurun_out/r02_multirank_parity_2gpu.jsonl | cut -c 1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_scale_2.json 2> gpurun_out/r02_scale_2.err; tail -2 gpurun_out/r02_scale_2.err
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_scale_2.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","n_gpus","allreduce")}, r["e2e"]["value"], r.get("full_sweep"), r["roofline"]["kernels"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(r['ms_per_step'], r['cpu_baseline'])"
