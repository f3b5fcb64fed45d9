"""Worker of tests/test_gpu_multirank.py: one process per rank (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* in the
environment), every rank drives its own BornEngine through libbm.so on its sample shard.

With >= WORLD_SIZE GPUs each rank owns one GPU and the control plane is NCCL (both all-reduce flavours run); on a
single-GPU box all ranks share cuda:0 (time-sliced contexts), the control plane is gloo and only the
torch.distributed all-reduce is exercised (the fused peer kernel spins on its peers' flags and needs them co-resident).

modes:  teacher   before every step a standalone single-GPU engine is loaded with the sharded run's tensors; both take
                  the step: chi, singular values, sum ln|psi| and Z must agree; replicas stay bitwise identical.
        bas       FREE-RUNNING bars-and-stripes training (well conditioned: no near-zero psi), sharded vs single:
                  the NLL curves must agree (TNutils.py:1570-1631 learning_epoch_cached)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mps_born_machines_b200 import BornEngine, mpslib, sweep_schedule   # noqa: E402
from mps_born_machines_b200.parallel import shard_bounds, replicated_checksum   # noqa: E402


def main():
    mode, allreduce = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    ngpu = torch.cuda.device_count()
    local = rank if ngpu >= world else 0
    torch.cuda.set_device(local)
    backend = 'nccl' if ngpu >= world else 'gloo'
    kw = {'device_id': torch.device(f'cuda:{local}')} if backend == 'nccl' else {}
    dist.init_process_group(backend, **kw)
    res = {'mode': mode, 'allreduce': allreduce, 'backend': backend, 'world': world, 'gpus': ngpu}
    if mode == 'teacher':
        L, D, N = 24, 16, 1001                      # uneven shards on purpose
        imgs = mpslib.synthetic_images(N, L, seed=3, p_on=0.3)
        mps = mpslib.random_mps(L, 4, rng=np.random.default_rng(1), centre=0)
        lo, hi = shard_bounds(N, world, rank)
        eng = BornEngine(L, D, device=local)
        eng.set_mps(mps); eng.set_data(imgs[lo:hi])
        assert eng.N_total == N
        if allreduce == 'peer':
            assert eng.enable_peer_allreduce()
        single = None
        if rank == 0:
            single = BornEngine(L, D, device=local, standalone=True)
            single.set_data(imgs)
        worst_s = worst_ln = worst_z = 0.0
        going_right, ok = True, True
        for k in sweep_schedule(L):
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
        res.update(chi_equal=bool(ok), max_rel_s=worst_s, max_rel_sumlog=worst_ln, max_rel_Z=worst_z, replica_drift=drift,
                   launches=int(eng.launch_count()))
        eng.close()
    elif mode == 'bas':
        from oracle import born_oracle as o     # test infrastructure: the data set only
        imgs = o.bars_n_stripes_all(4)
        L, D, N = 16, 16, len(imgs)
        mps = mpslib.random_mps(L, 2, rng=np.random.default_rng(7), centre=0)
        lo, hi = shard_bounds(N, world, rank)
        eng = BornEngine(L, D, device=local)
        eng.set_mps(mps); eng.set_data(imgs[lo:hi])
        if allreduce == 'peer':
            assert eng.enable_peer_allreduce()
        single = BornEngine(L, D, device=local, standalone=True)
        single.set_mps(mps); single.set_data(imgs)
        curve, curve1 = [], []
        lr = 0.5
        for _ in range(16):
            eng.epoch(lr, 'l2', max_bond=D)
            single.epoch(lr, 'l2', max_bond=D)
            curve.append(eng.nll(imgs, 0)); curve1.append(single.nll(imgs, 0))
            lr *= 0.95
        drift = max(replicated_checksum(torch.from_numpy(t).cuda()) for t in eng.get_mps())
        res.update(nll_sharded=curve, nll_single=curve1, replica_drift=drift, bonds=eng.bond_sizes(), bonds_single=single.bond_sizes())
        eng.close(); single.close()
    if rank == 0:
        print('RESULT ' + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
