"""Fast kernel micro-benchmark: the three DMMA kernels at the c3 shapes (N samples, bond D) on a SHORT chain,
so set-up takes seconds.  Per-kernel device times come from the library's event timers (bm_set_timing)."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mps_born_machines_b200 import BornEngine, mpslib

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=60000)
ap.add_argument('--d', type=int, default=256)
ap.add_argument('--sites', type=int, default=24)
ap.add_argument('--steps', type=int, default=6)
ap.add_argument('--svd', default='gram')
ap.add_argument('--precision', default='fp64')
ap.add_argument('--check', action='store_true')
a = ap.parse_args()
L, D, N = a.sites, a.d, a.n
imgs = mpslib.synthetic_images(N, L)
k0 = L // 2 - a.steps // 2 - 2
rng = np.random.default_rng(11)
mats = []
for k in range(L):      # bonds: min(D, 2^k, 2^(L-k)) would be tiny at the ends; force D in the middle region
    dl = 1 if k == 0 else D
    dr = 1 if k == L - 1 else D
    mats.append(rng.standard_normal((dl, dr, 2)) / np.sqrt(dr * 2))
mps = mpslib.canonize(mpslib.squeeze_chain(mats), k0)
mps[k0] = mps[k0] / np.linalg.norm(mps[k0])
eng = BornEngine(L, D, svd=a.svd, precision=a.precision)
eng.set_data(imgs); eng.set_mps(mps); eng.env_build(k0)
k = k0
if a.check:
    ref = BornEngine(L, D, svd=a.svd, precision='fp64'); ref.set_data(imgs); ref.set_mps(mps); ref.env_build(k0)
    S1 = eng.grad_to_ref(k0, eng.step_grad(k0)); S0 = ref.grad_to_ref(k0, ref.step_grad(k0))
    print(json.dumps({'check': 'tf32 vs fp64 gradient at D=%d N=%d' % (D, N), 'max_abs_diff_over_max': float(np.abs(S1 - S0).max() / np.abs(S0).max()), 'rel_fro': float(np.linalg.norm(S1 - S0) / np.linalg.norm(S0))}))
    ref.close()
for _ in range(2):
    eng.step(k, 1e-3, True, renorm='l2', max_bond=D); k += 1
eng.set_timing(True)
for _ in range(a.steps):
    eng.step(k, 1e-3, True, renorm='l2', max_bond=D); k += 1
t = eng.get_timing()
out = {}
for name in ('env_advance', 'psi', 'grad'):
    ms, cnt = t[name]
    out[name] = {'ms': round(ms / cnt, 4), 'tflops': round(2.0 * D * D * N / (ms / cnt * 1e-3) / 1e12, 2)}
for name in ('grad_reduce', 'merge', 'svd'):
    ms, cnt = t[name]
    out[name] = {'ms': round(ms / cnt, 4)}
print(json.dumps({'n': N, 'D': D, 'kernels': out}))
