"""per-step timing of the c3 bench steps with the Gram level count and the host-side wall time per step."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from mps_born_machines_b200 import BornEngine, mpslib
N, L, D = int(os.environ.get('N', 60000)), 784, 256
k0, warm, steps = 384, 3, 10
imgs = mpslib.synthetic_images(N, L)
start = k0 - warm
mps = mpslib.random_mps(L, D, d=2, rng=np.random.default_rng(11), centre=start)
eng = BornEngine(L, D, svd=os.environ.get('SVD', 'gram'))
eng.set_data(imgs); eng.set_mps(mps); eng.env_build(start)
torch.cuda.synchronize()
k = start
eng.set_timing(True)
prev = eng.get_timing()
for i in range(warm + steps):
    t0 = time.perf_counter()
    out = eng.step(k, 1e-3, True, renorm='l2', max_bond=D, cutoff=1e-10, want_s=True); k += 1
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    cur = eng.get_timing()
    svd_ms = cur['svd'][0] - prev['svd'][0]
    prev = cur
    eng.eig_stats()
    s = out['s']
    print(json.dumps({'step': i, 'wall_ms': round(dt, 3), 'svd_ms': round(svd_ms, 3), 'levels': eng.last_gram_levels, 'chi': out['chi'],
                      's_last/s0': float(s[-1] / s[0]), 'n_above_3e-4': int((s > 3.16e-4 * s[0]).sum())}), flush=True)
