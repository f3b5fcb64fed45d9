mkdir -p gpurun_out
BM_EIG_DEBUG=1 python tools/eig_one.py 2>&1 | grep -E "sytrd phases|stage_ms" | tail -3
python tools/test_eig.py > gpurun_out/r02_eig_wy.jsonl 2>gpurun_out/r02_eig_wy.err; tail -2 gpurun_out/r02_eig_wy.err
python - <<PY
import json
for l in open("gpurun_out/r02_eig_wy.jsonl"):
    r=json.loads(l); print(r["case"], "lam %.1e ortho0 %.1e ortho %.1e resid %.1e ms %.3f"%(r["lam_err"],r["ortho_before_polish"],r["ortho"],r["resid"],r["ms_on_chip"]), r["stage_ms"])
PY
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
