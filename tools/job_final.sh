mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2>gpurun_out/r02_bench_final.err; echo rc=$?; tail -3 gpurun_out/r02_bench_final.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2>/dev/null; echo rc=$?
