import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mps_born_machines_b200 as bm
rng = np.random.default_rng(4242)
eng = bm.BornEngine(4, 256)
for n, m in ((2, 1), (2, 2), (3, 1), (3, 2), (5, 1), (5, 2), (9, 1), (9, 4)):
    Y = rng.standard_normal((n + 3, n)); G = Y.T @ Y
    lam, U, info = eng.eig_sym(G, m, on_chip=True)
    lr = np.linalg.eigvalsh(G)
    print(n, m, 'lam err %.1e' % np.abs(lam - lr).max(), 'ortho %.1e' % np.abs(U.T @ U - np.eye(m)).max(), 'resid %.1e' % np.abs(G @ U - U * lam[::-1][:m]).max(), info['ortho_dev'], U[:, 0])
eng.close()
