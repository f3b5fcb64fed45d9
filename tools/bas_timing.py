"""Config 1 (bars-and-stripes 4x4, 30 patterns, Dmax=16, 16 epochs): engine vs the numpy oracle, wall-clock."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import born_oracle as o
from mps_born_machines_b200 import BornEngine
imgs = o.bars_n_stripes_all(4).astype(np.int64)
mps0 = o.canonize(o.random_mps(16, 2, rng=np.random.default_rng(1)), 0)
t = time.perf_counter()
cost_o, _ = o.learning_epoch_cached([x.copy() for x in mps0], imgs, 16, 0.5, batch_size=30, max_bond=16, lr_update=lambda l: l * 0.95, gauge='fixed')
t_cpu = time.perf_counter() - t
eng = BornEngine(16, 16); eng.set_mps(mps0); eng.set_data(imgs)
eng.epoch(0.5, 'max', max_bond=16)           # warm-up epoch on a copy of the state
eng.set_mps(mps0)
import torch
torch.cuda.synchronize(); t = time.perf_counter()
lr = 0.5; cost = []
for _ in range(16):
    eng.epoch(lr, 'max', max_bond=16); cost.append(eng.nll(imgs, 0)); lr *= 0.95
torch.cuda.synchronize(); t_gpu = time.perf_counter() - t
eng.set_timing(True); eng.epoch(lr, 'max', max_bond=16); tm = eng.get_timing()
print(json.dumps({'config': 'c1 BAS 4x4, 16 epochs = 480 steps', 'cpu_oracle_s': round(t_cpu, 3), 'gpu_engine_s': round(t_gpu, 3),
                  'gpu_ms_per_step': round(1e3 * t_gpu / 480, 3), 'final_nll': [cost_o[-1], cost[-1]],
                  'device_ms_per_call': {k: round(v[0] / max(1, v[1]), 4) for k, v in tm.items()}}))
