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
eng = BornEngine(L, D, svd=a.svd)
eng.set_data(imgs); eng.set_mps(mps); eng.env_build(k0)
k = k0
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
