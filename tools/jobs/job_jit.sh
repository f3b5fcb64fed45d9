mkdir -p gpurun_out
for cfg in "10" "10" "100" "100" "0"; do
  if [ "$cfg" = "0" ]; then X="--clock-ms 100000"; else X="--clock-ms $cfg"; fi
  timeout 200 python bench.py --steps 60 --warmup 3 --no-cpu --no-others --no-sweep --no-sampler --per-step $X > gpurun_out/r02_jit.json 2>gpurun_out/r02_jit.err
  python - <<PY
import json
r=json.loads(open("gpurun_out/r02_jit.json").read().strip().splitlines()[-1])
s=sorted(r["step_ms"]); print("$cfg", r["ms_per_step"], "median", s[len(s)//2], "max3", s[-3:], "n>2.6:", sum(x>2.6 for x in s))
PY
done
