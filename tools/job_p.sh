mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tf32" 2>&1 | tail -15
for p in tf32 tf32_reg; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-others --no-sweep --no-sampler --precision $p > gpurun_out/r02_bench_$p.json 2>gpurun_out/r02_bench_$p.err; tail -2 gpurun_out/r02_bench_$p.err
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_bench_$p.json").read().strip().splitlines()[-1])
print("$p", {k:r[k] for k in ("value","ms_per_step")}, r["roofline"]["kernels"]["grad"], r["step_ms"][:5])
PY
done
