#!/bin/bash
# ncu evidence for round 2 (run under gpurun, one GPU): plain run first (must exit 0), then the launch list of the same
# command, then --set full captures of the three DMMA kernels and of the tridiagonalisation.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others"
timeout 280 $CMD > gpurun_out/r02_prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_prof_plain.log; exit 1; }
tail -c 300 gpurun_out/r02_prof_plain.log; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1200 --csv --log-file gpurun_out/r02_launches_c3.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_c3.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_psi|k_grad|k_env_advance' -s 790 -c 7 -o gpurun_out/r02_prof_dmma $CMD > gpurun_out/r02_ncu_full1.log 2>&1
echo "full dmma rc=$?"
timeout 120 python tools/eig_one.py > gpurun_out/r02_eig_one_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sytrd_cluster|k_tri_eigvals|k_tri_vectors3|k_wy_apply|k_wy_T' -s 5 -c 5 -o gpurun_out/r02_prof_eig python tools/eig_one.py > gpurun_out/r02_ncu_full2.log 2>&1
echo "full eig rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
