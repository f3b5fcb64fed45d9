BM_EIG_DEBUG=3 timeout 100 python tools/eig_small.py 128 2>&1 | grep -E "tail phases|stage_ms" | tail -3
