L=784 D=256 BM_EIG_DEBUG=1 timeout 300 python tools/dbg_steep.py 2>&1 | grep -v "eig ticks" | tail -20
