mkdir -p gpurun_out
N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_scale_$N.json 2> gpurun_out/r02_scale_$N.err; echo rc=$?
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_scale_$N.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","n_gpus","allreduce")}, r["e2e"]["value"], r.get("full_sweep"), r["roofline"]["kernels"])
PY
