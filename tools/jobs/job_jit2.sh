mkdir -p gpurun_out
for mode in 0 1 0 1; do
  BM_STEP_SYNC=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 60 --warmup 3 --no-cpu --no-others --no-sweep --no-sampler > gpurun_out/r02_jit2.json 2>gpurun_out/r02_jit2.err
  python - <<PY
import json
r=json.loads(open("gpurun_out/r02_jit2.json").read().strip().splitlines()[-1])
s=sorted(r["step_ms"]); print("sync=$mode", round(r["ms_per_step"],3), "median", s[len(s)//2], "max4", s[-4:], "n>2.6:", sum(x>2.6 for x in s), "levels>1:", sum(l>1 for l in r["gram_levels"]), [i for i,x in enumerate(r["step_ms"]) if x>2.6])
PY
done
