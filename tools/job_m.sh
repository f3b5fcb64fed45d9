BM_EIG_DEBUG=1 python bench.py --steps 20 --warmup 3 --no-cpu --no-sampler --no-sweep --no-others 2>&1 | grep -E "bm split" | head
