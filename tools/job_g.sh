mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_g.log; cat gpurun_out/r02_pytest_g.log
cat gpurun_out/multirank_parity.jsonl
python tools/prof_facade.py 2>&1 | sed -n 1,12p
