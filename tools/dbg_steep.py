import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mps_born_machines_b200 import BornEngine, mpslib
from oracle import born_oracle as o
L, D, N = int(os.environ.get('L', 40)), int(os.environ.get('D', 64)), 2000
for decades in (4.0, 10.0):
    mps = mpslib.random_mps(L, D, rng=np.random.default_rng(11), centre=L // 2, decades=decades)
    imgs = mpslib.synthetic_images(N, L)
    k = L // 2
    A = o.merge(mps, k)
    s = np.linalg.svd(A.transpose(0, 1, 2, 3).reshape(A.shape[0] * A.shape[1], -1), compute_uv=False)
    print('decades', decades, 'merged singular values', s[0], s[D // 2], s[D - 1], s[min(len(s) - 1, D + 5)], 'finite', np.isfinite(A).all(), flush=True)
    for svd in ('gram', 'gram_syevd'):
        eng = BornEngine(L, D, svd=svd)
        eng.set_mps(mps); eng.set_data(imgs)
        try:
            for kk in range(k, k + 3):
                out = eng.step(kk, 1e-3, True, 'l2', max_bond=D, want_s=True)
                eng.eig_stats()
                print(svd, 'bond', kk, 'chi', out['chi'], 'levels', eng.last_gram_levels, 's', out['s'][0], out['s'][-1], 'Z', out['Z'], 'nzero', out['n_zero_psi'], eng.eig_stats(), eng.eig_fallback_reasons(), flush=True)
        except Exception as e:
            print(svd, 'FAILED', e, eng.eig_fallback_reasons(), flush=True)
        eng.close()
