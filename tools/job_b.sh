mkdir -p gpurun_out
python tools/dbg_c3left.py > gpurun_out/r02_dbg_c3left.log 2>&1; cat gpurun_out/r02_dbg_c3left.log | tail -30
