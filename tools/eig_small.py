import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import mps_born_machines_b200 as bm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rng = np.random.default_rng(0)
Y = rng.standard_normal((n + 5, n))
eng = bm.BornEngine(4, 256)
for _ in range(2):
    lam, U, info = eng.eig_sym(Y.T @ Y, n // 2, on_chip=True)
print(info)
eng.close()
