#!/bin/bash
# launch list (per-launch durations) of the final round-2 build: one ncu pass, after the same command exited 0 without ncu
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others"
timeout 280 $CMD > gpurun_out/r02e_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02e_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1200 --csv --log-file gpurun_out/r02e_launches_c3.csv $CMD > gpurun_out/r02e_ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02e_launches_c3.csv
