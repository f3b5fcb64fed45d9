timeout 400 python -m pytest tests -x -q -m gpu -k "reconstruct or gram or split or compress or eigen" 2>&1 | tail -3
timeout 600 python - <<PY
import bench, argparse, json
a = argparse.Namespace(sampler_bond=500, warmup=1)
r = bench.reconstruct_record(a, 0)
print(json.dumps(r["top_half_unknown"]))
PY
