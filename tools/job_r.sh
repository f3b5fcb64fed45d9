timeout 600 python - <<PY
import bench, argparse, json
a = argparse.Namespace(sampler_bond=500, warmup=1)
print(json.dumps(bench.reconstruct_record(a, 0)))
PY
