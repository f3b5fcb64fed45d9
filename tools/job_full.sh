mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_full.json 2>gpurun_out/r02_bench_full.err; echo rc=$?; tail -3 gpurun_out/r02_bench_full.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; tail -c 400 gpurun_out/r02_bench_reference.json
