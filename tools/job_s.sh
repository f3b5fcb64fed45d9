mkdir -p gpurun_out
timeout 300 python bench.py --workload c3steep_lr1e-9 --steps 10 --warmup 3 --no-cpu --no-others --no-sweep --no-sampler > gpurun_out/r02_bench_c3steep.json 2> gpurun_out/r02_bench_c3steep.err; tail -5 gpurun_out/r02_bench_c3steep.err
python -c "
import json,sys
r=json.loads(open('gpurun_out/r02_bench_c3steep.json').read().strip().splitlines()[-1]); print(r['ms_per_step'], r['step_ms'], r['gram_levels'], r['chi'], r['split_eig'], r['roofline']['kernels']['svd'])"
