mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -k "sampl or reconstruct or generate" 2>&1 | tail -5
timeout 200 python bench.py --workload c4 --steps 1 --warmup 1 > gpurun_out/r02_bench_c4.json 2>gpurun_out/r02_bench_c4.err; tail -2 gpurun_out/r02_bench_c4.err
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_bench_c4.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}, r["e2e"]["value"], r["roofline"])
PY
BM_SAMPLE_FUSED=1 timeout 200 python bench.py --workload c4 --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fused', r['value'], r['ms_per_step'])"
