mkdir -p gpurun_out
BM_EIG_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-sweep --no-others --no-cpu > gpurun_out/r02_scale_2dbg.json 2> gpurun_out/r02_scale_2dbg.err; echo rc=$?
grep -c "bm eig ticks" gpurun_out/r02_scale_2dbg.err
python - <<PY
import json
r=json.loads(open("gpurun_out/r02_scale_2dbg.json").read().strip().splitlines()[-1])
print(r["step_ms"], r["gram_levels"])
PY
