#!/bin/bash
# final ncu pass of round 2: launch list of the final build + full capture of the tridiagonalisation kernels
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others"
timeout 280 $CMD > gpurun_out/r02c_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02c_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1200 --csv --log-file gpurun_out/r02c_launches_c3.csv $CMD > gpurun_out/r02c_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 120 python tools/eig_one.py > gpurun_out/r02c_eig_one_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sytrd_cluster|k_sytrd_tail' -s 2 -c 2 -o gpurun_out/r02c_prof_sytrd python tools/eig_one.py > gpurun_out/r02c_ncu_full.log 2>&1
echo "full sytrd rc=$?"
ls -la gpurun_out/r02c*
