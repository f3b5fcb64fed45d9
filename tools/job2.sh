mkdir -p gpurun_out
BM_EIG_DEBUG=1 python tools/eig_one.py 2>&1 | grep -E "sytrd phases|sub-phases|stage_ms" | tail -3
python tools/test_eig.py > gpurun_out/r02_eig_v3.jsonl 2>gpurun_out/r02_eig_v3.err; tail -2 gpurun_out/r02_eig_v3.err
python - <<PY
import json
for l in open("gpurun_out/r02_eig_v3.jsonl"):
    r=json.loads(l); print(r["case"], "lam %.1e ortho0 %.1e ortho %.1e resid %.1e ms %.3f"%(r["lam_err"],r["ortho_before_polish"],r["ortho"],r["resid"],r["ms_on_chip"]), r["stage_ms"])
PY
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu --no-sampler > gpurun_out/r02_bench_v3.json 2>gpurun_out/r02_bench_v3.err; tail -3 gpurun_out/r02_bench_v3.err; python - <<PY
import json
r=json.loads(open("gpurun_out/r02_bench_v3.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","step_ms","gram_levels","gpu_launches")}, r["e2e"], r["roofline"]["kernels"])
PY
