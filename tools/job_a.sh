mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_a.log; cat gpurun_out/r02_pytest_a.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a.json 2>gpurun_out/r02_bench_a.err; tail -3 gpurun_out/r02_bench_a.err; cat gpurun_out/r02_bench_a.json
