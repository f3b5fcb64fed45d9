#!/bin/bash
# ncu evidence for round 1 (run under gpurun, one GPU).  Plain run first (must exit 0), then the launch list of
# the same command, then one --set full capture of the three hot kernels.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 1 --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
tail -c 600 gpurun_out/prof_plain.log
# launch list: skip the 783 set-up advances + uploads, take the kernels of the warm-up + timed steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:'k_psi|k_grad$|k_grad\(' -c 4 -o gpurun_out/prof_psi_grad $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_env_advance -s 800 -c 2 -o gpurun_out/prof_env $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full2 rc=$?"
ls -la gpurun_out/
