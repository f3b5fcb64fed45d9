"""Host wall-clock breakdown of one two-site step (with a device sync after each phase) vs per-kernel device times."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mps_born_machines_b200 import BornEngine, mpslib
L, D, N = 64, 256, 60000
imgs = mpslib.synthetic_images(N, L)
k0 = 8
rng = np.random.default_rng(11)
mats = [rng.standard_normal((1 if k == 0 else D, 1 if k == L - 1 else D, 2)) / np.sqrt(2 * D) for k in range(L)]
mps = mpslib.canonize(mpslib.squeeze_chain(mats), k0); mps[k0] /= np.linalg.norm(mps[k0])
eng = BornEngine(L, D); eng.set_data(imgs); eng.set_mps(mps); eng.env_build(k0)
k = k0
for _ in range(2):
    eng.step(k, 1e-3, True, renorm='l2', max_bond=D); k += 1
sync = torch.cuda.synchronize
acc = {'grad': 0, 'update': 0, 'advance': 0, 'total_nosync': 0}
n = 6
for _ in range(n):
    sync(); t0 = time.perf_counter()
    S = eng.step_grad(k); sync(); t1 = time.perf_counter()
    eng.step_update(k, S, N, 1e-3, True, 'l2', D, 1e-10); sync(); t2 = time.perf_counter()
    eng.env_advance(k, True); sync(); t3 = time.perf_counter()
    acc['grad'] += t1 - t0; acc['update'] += t2 - t1; acc['advance'] += t3 - t2; k += 1
per = []
for _ in range(n):
    sync(); t0 = time.perf_counter(); eng.step(k, 1e-3, True, renorm='l2', max_bond=D); sync(); dt = time.perf_counter() - t0; acc['total_nosync'] += dt; per.append(round(dt * 1e3, 2)); k += 1
per2 = []
for _ in range(n):
    sync(); t0 = time.perf_counter()
    S = eng.step_grad(k); eng.step_update(k, S, N, 1e-3, True, 'l2', D, 1e-10); eng.env_advance(k, True); sync(); per2.append(round((time.perf_counter() - t0) * 1e3, 2)); k += 1
print('eng.step per-step ms', per, ' raw 3-call per-step ms', per2)
eng.set_timing(True)
for _ in range(n):
    eng.step(k, 1e-3, True, renorm='l2', max_bond=D); k += 1
t = eng.get_timing()
print(json.dumps({'host_ms': {a: round(1e3 * b / n, 3) for a, b in acc.items()}, 'device_ms': {a: round(b[0] / max(1, b[1]), 3) for a, b in t.items()}}))
