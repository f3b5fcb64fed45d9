#!/bin/bash
# second ncu pass of round 2: k_grad after the grid reorder, and the TF32 gradient kernels (TMA-staged and register-staged)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others"
timeout 280 $CMD > gpurun_out/r02b_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02b_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grad' -s 0 -c 4 -o gpurun_out/r02b_prof_grad $CMD > gpurun_out/r02b_ncu1.log 2>&1
echo "grad rc=$?"
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others --precision tf32"
timeout 280 $CMD2 > gpurun_out/r02b_plain2.log 2>&1 || { echo "plain tf32 run failed"; tail -5 gpurun_out/r02b_plain2.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grad_tf32_tma|k_pack_tf32' -s 0 -c 4 -o gpurun_out/r02b_prof_tma $CMD2 > gpurun_out/r02b_ncu2.log 2>&1
echo "tma rc=$?"
CMD3="python bench.py --steps 3 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others --precision tf32_reg"
timeout 280 $CMD3 > gpurun_out/r02b_plain3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grad_tf32' -s 0 -c 2 -o gpurun_out/r02b_prof_tf32reg $CMD3 > gpurun_out/r02b_ncu3.log 2>&1
echo "reg rc=$?"
ls -la gpurun_out/r02b*.ncu-rep
