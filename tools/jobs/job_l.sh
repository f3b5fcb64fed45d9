mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_pytest_l.log; cat gpurun_out/r02_pytest_l.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others > gpurun_out/r02_bench_l.json 2>gpurun_out/r02_bench_l.err; tail -3 gpurun_out/r02_bench_l.err
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_bench_l.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}, r["e2e"]["value"], r["e2e_engine_api"]["value"], r.get("full_sweep"), r["roofline"]["kernels"], r["roofline"]["traffic"], r["step_ms"], r["split_eig"])
PY
