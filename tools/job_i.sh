python tools/dbg_eig_small.py 2>&1 | tail -12
BM_BACKTRANSFORM=warp python tools/dbg_eig_small.py 2>&1 | tail -9
