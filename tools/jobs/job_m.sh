mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_pytest_m.log; cat gpurun_out/r02_pytest_m.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others --no-sweep > gpurun_out/r02_bench_m.json 2>gpurun_out/r02_bench_m.err; tail -3 gpurun_out/r02_bench_m.err
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_bench_m.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}, r["roofline"]["kernels"], r.get("sampler"))
PY
