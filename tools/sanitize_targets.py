"""Small invocations of the three kernels with hand-rolled synchronisation, for compute-sanitizer:
   eig   k_sytrd_cluster (DSMEM bulk copies + mbarrier parity) + the rest of the on-chip eigensolver, n = 96
   tf32  k_grad_tf32 (mbarrier / tcgen05.commit pipeline) inside two training steps, D = 32
   step  a plain FP64 two-site sweep (k_psi, k_grad, k_env_advance, k_dgemm, split), L = 10, D = 8"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mps_born_machines_b200 as bm
from mps_born_machines_b200 import mpslib

what = sys.argv[1]
if what == 'eig':
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((96, 96))
    eng = bm.BornEngine(4, 64)
    lam, U, info = eng.eig_sym(Y.T @ Y, 40, on_chip=True)
    ref = np.linalg.eigvalsh(Y.T @ Y)
    assert np.allclose(lam, ref, rtol=1e-10, atol=1e-10 * ref[-1]), 'eigenvalues'
    print('eig ok', info)
    eng.close()
else:
    L, D, N = 10, (32 if what == 'tf32' else 8), 300
    imgs = mpslib.synthetic_images(N, L, seed=3, p_on=0.3)
    mps = mpslib.random_mps(L, D, rng=np.random.default_rng(1), centre=3)
    eng = bm.BornEngine(L, D, precision='tf32' if what == 'tf32' else 'fp64')
    eng.set_mps(mps); eng.set_data(imgs)
    for k in (3, 4, 5):
        out = eng.step(k, 1e-3, True, 'l2', max_bond=D)
    print(what, 'ok chi', out['chi'], 'eig', eng.eig_stats())
    eng.close()
