mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for tgt in eig tf32 step; do
    echo "=== $tool $tgt" >> gpurun_out/r02_sanitizer.log
    timeout 600 $S --tool $tool --print-limit 5 python tools/sanitize_targets.py $tgt 2>&1 | grep -v "^$" | tail -12 >> gpurun_out/r02_sanitizer.log
  done
done
cat gpurun_out/r02_sanitizer.log
