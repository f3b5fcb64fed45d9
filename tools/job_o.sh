mkdir -p gpurun_out
BENCH_TRACE=1 BENCH_TRACE_DUMP_S=200 timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_scale_2.json 2> gpurun_out/r02_scale_2.err; echo rc=$?
grep -E "bench rank|File|line" gpurun_out/r02_scale_2.err | tail -40
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_scale_2.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","n_gpus","allreduce")}, r["e2e"]["value"], r.get("full_sweep"), r["roofline"]["kernels"])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(r['ms_per_step'], r['cpu_baseline'])"
