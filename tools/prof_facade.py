"""cProfile of the facade's per-step host path (learning_step_cached + write-back + sequential_update) at D=256."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mps_born_machines_b200 import BornEngine, mpslib, TNutils as T
L, D, N, k0 = 784, 256, 2000, 380
imgs = mpslib.synthetic_images(N, L)
mps = mpslib.random_mps(L, D, rng=np.random.default_rng(1), centre=k0)
eng = BornEngine(L, D); eng.set_data(imgs); eng.set_mps(mps); eng.env_build(k0)
fmps = T.MPS(eng.get_mps()); cache = T.ImgCache(eng, imgs)
kw = dict(renorm='l2', max_bond=D, cutoff=1e-10)
def loop(k, n):
    for _ in range(n):
        tn = T.learning_step_cached(fmps, k, None, 1e-3, cache, True, **kw)
        T._write_back(fmps, k, tn)
        T.sequential_update(fmps, None, cache, k, True)
        k += 1
    return k
k = loop(k0, 3)
pr = cProfile.Profile(); pr.enable(); k = loop(k, 20); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(25); print(s.getvalue())
