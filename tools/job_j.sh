BM_EIG_DEBUG=1 python bench.py --steps 6 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others 2>&1 | grep "eig ticks" | sed -n 4,9p
