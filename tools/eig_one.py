"""one on-chip eigen-decomposition at n=512, m=256 (ncu target)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import mps_born_machines_b200 as bm
rng = np.random.default_rng(0)
Y = rng.standard_normal((512, 512))
eng = bm.BornEngine(4, 256)
for _ in range(2):
    lam, U, info = eng.eig_sym(Y.T @ Y, 256, on_chip=True)
print(info)
eng.close()
