mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others --no-sweep > gpurun_out/r02_bench_m$i.json 2>gpurun_out/r02_bench_m$i.err
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_bench_m$i.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}, r["e2e"]["value"], r["e2e"]["host_wall_ms_per_step"], r["e2e_engine_api"]["value"])
PY
done
