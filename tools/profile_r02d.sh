#!/bin/bash
# ncu --set full capture of the three DMMA kernels of the final round-2 build (one ncu pass per gpurun call, after the
# same command exited 0 without ncu)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others"
timeout 280 $CMD > gpurun_out/r02d_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02d_plain.log; exit 1; }
tail -c 200 gpurun_out/r02d_plain.log; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_psi|k_grad|k_env_advance' -s 790 -c 7 -o gpurun_out/r02d_prof_dmma $CMD > gpurun_out/r02d_ncu_full.log 2>&1
echo "full dmma rc=$?"; ls -la gpurun_out/r02d*
