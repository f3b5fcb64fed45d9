"""accuracy of the level-1 singular values of the Gram split against the known spectrum, by position in the spectrum"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mps_born_machines_b200 as bm
D = 256
worst = {}
for seed in range(6):
    rng = np.random.default_rng(100 + seed)
    lo = [1e-9, 1e-8, 1e-7][seed % 3]          # log-uniform spectrum: the 256th of 512 values sits near sqrt(lo)
    s = np.sort(np.exp(rng.uniform(np.log(lo), 0.0, size=2 * D)))[::-1]; s[0] = 1.0
    U, _ = np.linalg.qr(rng.standard_normal((2 * D, 2 * D)))
    V, _ = np.linalg.qr(rng.standard_normal((2 * D, 2 * D)))
    # A = U diag(s) V^T as merged tensor of sites 1,2: m1 = (U s^(1/2))[:, :D]... use exact factorisation through a D-bond is impossible (rank 2D):
    # instead set site1 = U diag(s) (2D x 2D needs bond 2D): use an engine with cap 2D and bond 2D, then max_bond = D truncation
    m1 = (U * s).reshape(D, 2, 2 * D).transpose(0, 2, 1)
    m2 = V.T.reshape(2 * D, D, 2)
    mps = [rng.standard_normal((D, 2)), m1, m2, rng.standard_normal((D, 2))]
    eng = bm.BornEngine(4, 2 * D)
    eng.set_mps(mps)
    for gr in (True, False):
        out = eng.split_bond(1, going_right=gr, max_bond=D, cutoff=1e-10, want_s=True)
        eng.set_mps(mps)
        rel = np.abs(out['s'] - s[:D]) / s[:D]
        eng.eig_stats()
        key = f'seed{seed}_{"r" if gr else "l"}'
        worst[key] = {'levels': eng.last_gram_levels, 'max_rel': float(rel.max()), 'at_ratio': float(s[int(rel.argmax())]), 's_D': float(s[D - 1]),
                      'max_rel_below_3e-4': float(rel[s[:D] < 3.2e-4].max()) if (s[:D] < 3.2e-4).any() else 0.0}
    eng.close()
print(json.dumps({'tau': os.environ.get('BM_GRAM_TAU', '1e-7'), 'levels': [v['levels'] for v in worst.values()], 's_D': ['%.1e' % v['s_D'] for v in worst.values()],
                  'max_rel': ['%.1e' % v['max_rel'] for v in worst.values()], 'at': ['%.1e' % v['at_ratio'] for v in worst.values()]}))
