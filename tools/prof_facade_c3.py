"""Host-side timeline of the facade's per-step path at the c3 shape (N=60000, D=256): where the e2e step spends its time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mps_born_machines_b200 import BornEngine, mpslib, TNutils as T
L, D, N, k0 = 784, 256, 60000, 380
imgs = mpslib.synthetic_images(N, L)
mps = mpslib.random_mps(L, D, rng=np.random.default_rng(1), centre=k0)
eng = BornEngine(L, D); eng.set_data(imgs); eng.set_mps(mps); eng.env_build(k0)
fmps = T.MPS(eng.get_mps()); cache = T.ImgCache(eng, imgs)
kw = dict(renorm='l2', max_bond=D, cutoff=1e-10)
acc = {}
def wrap(name):
    f = getattr(eng, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(eng, name, g)
for n in ('set_site', 'get_site', 'step', 'env_build', 'env_advance'):
    wrap(n)
def loop(k, n):
    for _ in range(n):
        tn = T.learning_step_cached(fmps, k, None, 1e-3, cache, True, **kw)
        T._write_back(fmps, k, tn)
        T.sequential_update(fmps, None, cache, k, True)
        k += 1
    return k
k = loop(k0, 5)
acc.clear()
torch.cuda.synchronize(); t0 = time.perf_counter(); n = 40
k = loop(k, n)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print(f'per step {tot / n * 1e3:.3f} ms;', {a: round(b / n * 1e3, 3) for a, b in acc.items()}, 'other', round((tot - sum(acc.values())) / n * 1e3, 3))
# the same loop through the engine API with pinned buffers
