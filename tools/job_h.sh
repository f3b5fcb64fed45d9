mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_h.log; cat gpurun_out/r02_pytest_h.log
