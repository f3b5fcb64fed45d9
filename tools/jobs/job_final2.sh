mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_pytest_final.log; cat gpurun_out/r02_pytest_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2>gpurun_out/r02_bench_final.err; echo rc=$?; tail -3 gpurun_out/r02_bench_final.err
