mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_e.json 2>gpurun_out/r02_bench_e.err; echo rc=$?; tail -5 gpurun_out/r02_bench_e.err
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_bench_e.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}, r["e2e"], r["e2e_engine_api"], r.get("full_sweep"), r["roofline"]["peak"], r["roofline"]["frac"], r["roofline"]["whole_step_frac"])
print(json.dumps(r.get("other_configs"))[:3000])
print(r.get("cpu_baseline"))
PY
