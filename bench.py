#!/usr/bin/env python
"""bench.py — headline benchmark of the MPS Born machine hot path on B200.

Metric (BASELINE.json): sample·site updates / s during the two-site sweep at Dmax=256.
One "step" = one two-site step of the sweep over ALL training samples at a mid-chain bond with full bond
dimension: merge, psi_n, S = sum psi'_n/psi_n, (all-reduce), update + renormalise, SVD split, write-back,
one-site environment advance  (learning_step_cached + sequential_update, TNutils.py:1510-1568, :504-526).
One sample·site update = one sample's share of one step.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (numpy oracle port)

Workload "c3" (default): synthetic binarised 28x28, 60 000 samples in total (sharded over the N GPUs:
strong scaling), L=784, D=256 everywhere (random right/left-canonical MPS, centre at the first timed bond).
"""
import argparse
import gc
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

# torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the reference arm (numpy on the host cores, rank 0
# only) must use all of them, so its BLAS/OpenMP pools are sized BEFORE numpy is imported (the N>1
# reference lines ran on one thread).  Our own arm keeps the launcher's setting: N ranks with a full-size pool each
# oversubscribe the host during set-up (the QR sweeps of the random MPS took minutes instead of seconds).
if '--impl' in sys.argv and 'reference' in sys.argv and os.environ.get('RANK', '0') == '0':
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N_total, L, D, first bond)
    'c3': (60000, 784, 256, 384),
    'c2': (1000, 784, 100, 384),
    'c5': (10000, 784, 128, 384),        # BASELINE config 5 as worded: 4-level pixels as one-hot d=4 sites
    'c5b': (10000, 1568, 128, 768),      # the reference's grayscale notebook: 2 bits per pixel, d stays 2, L doubles
    'tiny': (512, 64, 32, 24),
    # c3 shape with a Schmidt spectrum decaying over ~10 decades at every cut (stand-in for a trained model): the Gram
    # split then needs several deflation levels per step; fewer samples, the split is what is measured
    'c3steep': (8192, 784, 256, 384),
    'c3steep_lr1e-9': (8192, 784, 256, 384),     # the same with a vanishing learning rate: the update does not lift the tail
}
DECADES = {'c3steep': 10.0, 'c3steep_lr1e-9': 10.0}
LR = {'c3steep_lr1e-9': 1e-9}
PHYS_DIM = {'c5': 4}
FP64_PEAK_FALLBACK = 35.7  # cuBLAS DGEMM 8192^3 sustained, measured on this pool in round 1 (profiles/r01_probe_fp64_svd.jsonl);
                           # only used when the live probe below cannot run


def fp64_peak_live(device):
    """FP64 roofline denominator measured IN THIS RUN: sustained cuBLAS DGEMM (torch.matmul, float64, 8192^3) on the
    bench GPU, CUDA events, best of 3 after a warm-up.  FP64 is not in the driver-written MEASURED_PEAKS.json and
    tcgen05 has no FP64 kind, so the library DGEMM is the reference point for DMMA kernels."""
    import torch
    try:
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device=device)
        b = torch.randn(n, n, dtype=torch.float64, device=device)
        c = torch.empty(n, n, dtype=torch.float64, device=device)
        torch.matmul(a, b, out=c)
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        del a, b, c
        torch.cuda.empty_cache()
        return best, 'live: cuBLAS DGEMM 8192^3 (torch.matmul float64) timed in this run, best of 3'
    except Exception as exc:      # e.g. not enough free memory next to the environment stack
        return FP64_PEAK_FALLBACK, f'fallback 35.7 (round-1 probe; live probe failed: {exc})'


def ncu_traffic(kernel, workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    of this round (profiles/r02_ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when no
    capture exists for this exact kernel/workload/GPU count."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')) as f:
            tab = json.load(f)
        ent = tab.get(f'{workload}:{world}', {}).get(kernel)
        return (ent['dram_bytes'], ent.get('source')) if ent else (None, None)
    except Exception:
        return None, None


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(x.get('num_threads', 1)) for x in threadpool_info()] or [1])
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).
    The timed region of the headline workload is only ~60 ms, shorter than nvidia-smi's start-up, so the sampler is
    started (and known to be printing) BEFORE the warm-up steps; `mark()` brackets the timed region and `stop()`
    reports the samples whose timestamps fall inside it (falling back to everything since the start, with a note)."""
    Q = 'timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index=0, period_ms=10):
        self.period_ms = int(period_ms)
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        self.gpu = gpu_index
        self.t0 = self.t1 = None

    def start(self, wait_s=3.0):
        try:
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.gpu), '-lms', str(self.period_ms)], stdout=self.f, stderr=subprocess.DEVNULL)
            t_end = time.time() + wait_s
            while time.time() < t_end and os.path.getsize(self.path) == 0:     # first sample printed: sampler is live
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark(self):
        if self.t0 is None:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    @staticmethod
    def _ts(text):
        import datetime
        return datetime.datetime.strptime(text.strip(), '%Y/%m/%d %H:%M:%S.%f').timestamp()

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(',')]
                if len(p) < 9:
                    continue
                try:
                    ts = self._ts(p[0])
                except Exception:
                    ts = None
                try:
                    rows.append((ts, float(p[1]), float(p[2]),
                                 [name for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9])
                                  if v.lower().startswith('active')]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if r[0] is not None and self.t0 is not None and self.t1 is not None
                  and self.t0 - 0.015 <= r[0] <= self.t1 + 0.015]
        use, window = (inside, 'timed region') if inside else (rows, 'warm-up + timed region (no sample fell inside the timed region)')
        if use:
            out['sm_mhz'] = float(np.median([r[1] for r in use])); out['sm_max_mhz'] = float(np.max([r[2] for r in use]))
            out['reasons'] = sorted({x for r in use for x in r[3]})
            out['samples'] = len(use); out['window'] = window
        return out


# ------------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: numpy float64 oracle port on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_steps(N, D, steps, seed=5, d=2):
    """Time `steps` two-site steps (TNutils.py:1510-1568 arithmetic + sequential_update) of the oracle port
    at bond dimension D on N samples with all host BLAS threads.  Environments are synthetic random vectors
    (the arithmetic and its cost do not depend on their values); pixel statistics are those of the workload."""
    from oracle import born_oracle as o
    rng = np.random.default_rng(seed)
    imgs = o.synthetic_images(N, 4)           # 4 sites are enough: columns 1,2 play bonds k,k+1
    if d > 2:
        imgs = rng.integers(0, d, size=imgs.shape)
    phys = o.embed(imgs, d)
    # 4-site MPS with open bonds D in the middle: sites 1,2 are the pair being optimised
    mps3 = [rng.standard_normal((1, D, d)) / math.sqrt(d * D), rng.standard_normal((D, D, d)) / math.sqrt(d * D),
            rng.standard_normal((D, D, d)) / math.sqrt(d * D), rng.standard_normal((D, 1, d)) / math.sqrt(d * D)]
    mps = o.from_site3(mps3)
    lv = rng.standard_normal((N, D)); rv = rng.standard_normal((N, D))
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out = o.two_site_step(mps, 1, lv, rv, phys[:, 1], phys[:, 2], 1e-3, True, 'l2', max_bond=D)
        o.write_back(mps, 1, out['left'], out['right'])
        new = o._grouped_apply_right(lv, o.site3(mps, 1), phys[:, 1])      # sequential_update for all N
        new, _ = o._rescale(new, 'maxabs')
        times.append(time.perf_counter() - t0)
        lv = rng.standard_normal((N, D))                                   # fresh inputs (no cache reuse)
    return times


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    N, L, D, k0 = WORKLOADS[args.workload]
    cores = os.cpu_count()
    dd = PHYS_DIM.get(args.workload, 2)
    cpu_reference_steps(min(N, 2000), D, 1, d=dd)                 # warm BLAS
    t = []
    for _ in range(max(args.warmup, 0)):
        cpu_reference_steps(N, D, 1, d=dd)
    t = cpu_reference_steps(N, D, args.steps, d=dd)
    ms = 1e3 * float(np.mean(t))
    value = N / float(np.mean(t))
    line = {
        'impl': 'reference', 'metric': 'sample_site_updates_per_s_two_site_sweep', 'value': value,
        'unit': 'sample·site updates/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(args.workload, int(os.environ.get('WORLD_SIZE', args.gpus))),
        'cpu_baseline': {'value': value, 'unit': 'sample·site updates/s', 'cores': cores, 'kind': 'port',
                         'blas_threads': blas_threads(),
                         'sample': f'{args.steps} mid-chain two-site steps (psi+grad+SVD+env advance) at N={N}, D={D}, '
                                   f'numpy float64 oracle port, BLAS pool = {blas_threads()} threads on {cores} host cores '
                                   f'(OMP_NUM_THREADS reset from the launcher default before numpy was imported)'},
        'e2e': {'value': value, 'unit': 'sample·site updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(name, world):
    N, L, D, k0 = WORKLOADS[name]
    return {'workload': f'{name}: synthetic binarized 28x28 (L={L}), {N} samples total, Dmax={D}, ' + ('Schmidt spectra decaying over ~10 decades, ' if name in DECADES else '') + (f'lr={LR[name]:g}, ' if name in LR else '') +
                        f'mid-chain two-site steps from bond {k0} (full bond dimension), l2 renorm, cutoff 1e-10',
            'n_samples': N, 'sites': L, 'Dmax': D, 'parallelism': f'samples sharded over {world} GPU(s), '
            'one all-reduce of the 2 MiB gradient per step (fused into the split-reduction kernel over NVLink peer memory, or NCCL), replicated SVD' if world > 1 else 'single GPU',
            'l2': 'each step reads fresh environment slots (R slot untouched since set-up); per-step inputs '
                  '> L2 at N=1, L2 flushed between timed steps otherwise'}


# ------------------------------------------------------------------------------------------------------
def trace(msg):
    if os.environ.get('BENCH_TRACE'):
        print(f"[bench rank {os.environ.get('RANK', '0')} t={time.time() % 1000:.1f}] {msg}", file=sys.stderr, flush=True)


def measure_steps(args, workload, rank, world, local, full=True):
    """Set up `workload` on this rank's shard and time `args.steps` mid-chain two-site steps (device-resident inputs,
    CUDA events per step).  full=True adds the e2e loops, the per-kernel roofline pass, the full-chain sweep and the
    secondary sampler measurement (the headline line); full=False returns the compact record used for the other
    BASELINE configs."""
    import torch
    import torch.distributed as dist
    from mps_born_machines_b200 import BornEngine
    from mps_born_machines_b200 import mpslib

    N, L, D, k0 = WORKLOADS[workload]
    if args.samples and full:
        N = args.samples
    steps, warm = args.steps, args.warmup
    assert k0 + warm + 3 * steps + 16 < L - 2, 'not enough sites for the requested number of steps'

    # ---- set-up (untimed): data shard, MPS with the centre at the first bond, environments
    t_setup = time.time()
    d = PHYS_DIM.get(workload, 2)
    if d == 2:
        imgs = mpslib.synthetic_images(N, L)
    else:   # 4-level synthetic images: two independent bit planes of the prototype mixture
        imgs = 2 * mpslib.synthetic_images(N, L, seed=1234) + mpslib.synthetic_images(N, L, seed=4321, p_on=0.3)
    per = (N + world - 1) // world
    shard = imgs[rank * per:min(N, (rank + 1) * per)]
    start = k0 - warm
    mps = mpslib.random_mps(L, D, d=d, rng=np.random.default_rng(11), centre=start, decades=DECADES.get(workload, 0.0))
    eng = BornEngine(L, D, phys_dim=d, device=local, svd=args.svd, precision=args.precision)
    eng.set_data(shard, n_total=N)
    eng.set_mps(mps)
    allreduce_mode = 'none' if world == 1 else args.allreduce
    if world > 1 and args.allreduce == 'peer':
        try:
            eng.enable_peer_allreduce()
        except Exception as exc:          # e.g. CUDA IPC unavailable in this container: fall back to NCCL, say so
            allreduce_mode = f'nccl (peer-memory path unavailable: {exc})'
    trace('setup done')
    gc.collect(); gc.freeze()      # a full collection over torch's object graph costs tens of ms: not inside a timed loop
    eng.env_build(start)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup
    flush = None
    if shard.shape[0] * D * 8 * 2 < 2 * 126e6 or getattr(args, 'per_step', False):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=f'cuda:{local}')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=f'cuda:{local}')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    kw = dict(renorm='l2', max_bond=D, cutoff=1e-10)
    lr = LR.get(workload, 1e-3)
    k = start
    clocks = ClockSampler(local, getattr(args, 'clock_ms', 10))
    if rank == 0 and full:
        clocks.start()                    # live before the warm-up: the timed region is shorter than its start-up
    for _ in range(warm):
        eng.step(k, lr, True, **kw); k += 1

    trace('warm-up done')
    # ---- timed: EXACTLY `steps` steps, device-resident inputs.  With per-step inputs larger than L2 (N = 1) the K steps
    #      are ONE library call (BornEngine.sweep -> bm_sweep, what epoch() uses); otherwise L2 is flushed between the
    #      steps and each step is its own call.
    l0 = eng.launch_count()
    eig0 = eng.eig_stats()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    chis, levels = [], []
    barrier()
    clocks.mark()
    wall0 = time.perf_counter()
    if flush is None:
        ev[0][0].record()
        outs = eng.sweep(list(range(k, k + steps)), 1, lr, **kw); k += steps
        ev[0][1].record()
        chis = [o_['chi'] for o_ in outs]
        eng.eig_stats(); levels = [eng.last_gram_levels]
        ev = ev[:1]
    else:
        for i in range(steps):
            flush.fill_(i & 0xff)
            barrier()
            ev[i][0].record()
            out = eng.step(k, lr, True, **kw); k += 1
            ev[i][1].record()
            chis.append(out['chi'])
            eng.eig_stats(); levels.append(eng.last_gram_levels)
    barrier()
    wall = time.perf_counter() - wall0
    clocks.mark()
    step_ms = [round(a.elapsed_time(b), 3) for a, b in ev]      # one entry per library call (the whole run when sweep() was used)
    launches = eng.launch_count() - l0
    eig1 = eng.eig_stats()
    dev_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    clk = clocks.stop() if (rank == 0 and full) else {}
    value = N * steps / (dev_ms * 1e-3)
    rec = {'value': value, 'ms_per_step': dev_ms / steps, 'step_ms': step_ms, 'gram_levels': levels, 'chi': int(chis[-1]),
           'gpu_launches': int(launches), 'split_eig': {'on_chip': eig1[0] - eig0[0], 'cusolver_fallback': eig1[1] - eig0[1], 'fallback_reasons_since_setup': eng.eig_fallback_reasons()},
           'wall_s_timed': wall, 'setup_s': t_setup, 'allreduce': allreduce_mode, 'timed_call': 'sweep (one bm_sweep call)' if flush is None else 'step (L2 flushed between steps)', 'clocks': clk, 'N': N, 'L': L, 'D': D, 'd': d}
    if not full:
        eng.set_timing(True)
        for _ in range(3):
            eng.step(k, lr, True, **kw); k += 1
        tim = eng.get_timing(); eng.set_timing(False)
        rec['kernel_ms'] = {n: tim[n][0] / tim[n][1] for n in ('grad', 'psi', 'env_advance', 'merge', 'svd') if tim[n][1]}
        eng.close()
        return rec

    trace('timed region done')
    # ---- replicas: every rank carries the whole MPS and repeats the split on the all-reduced gradient; after the timed
    #      steps the site tensors must be IDENTICAL on all ranks (checksum spread 0.0).  Untimed.
    try:
        cs, nsite = 0.0, 0
        for j in range(start, min(k + 2, L)):
            t = eng.get_site(j).ravel()
            cs += float(np.dot(t, np.cos(np.arange(t.size, dtype=np.float64))))
            nsite += 1
        hi, lo = max_over_ranks(cs), -max_over_ranks(-cs)
        rec['multi_gpu_parity'] = {'replica_checksum_spread': hi - lo, 'replicas_identical': bool(hi == lo), 'sites_compared': nsite,
                                   'ranks': world, 'teacher_forced_vs_single_gpu': 'tests/test_gpu_multirank.py (profiles/r02_multirank_parity_2gpu.jsonl)'}
    except Exception as exc:                  # never lose the bench line over the side check
        rec['multi_gpu_parity'] = {'error': str(exc)[:200]}
    # ---- e2e (a): the reference's own call sequence through the facade with HOST tensors, every step:
    #      learning_step_cached -> write the two tensors back into the host MPS -> sequential_update
    #      (TNutils.py:1603-1620).  Uploads the site tensor the engine has not produced itself and downloads 2 per step (pageable numpy memory, as a
    #      notebook user would hold it).
    from mps_born_machines_b200 import TNutils as T
    fmps = T.MPS(eng.get_mps())
    cache = T.ImgCache(eng, shard)
    site_bytes = lambda j: int(fmps.tensors[j].data.nbytes)
    h2d = d2h = 0
    for i in range(min(args.warmup, 3)):   # warm-up of the facade path itself (staging buffers, first foreign uploads)
        tn = T.learning_step_cached(fmps, k, None, lr, cache, True, **kw)
        T._write_back(fmps, k, tn)
        T.sequential_update(fmps, None, cache, k, True)
        k += 1
    eng._h2d_bytes = 0
    barrier()
    e0 = time.perf_counter()
    wall = []
    for i in range(steps):
        w0 = time.perf_counter()
        tn = T.learning_step_cached(fmps, k, None, lr, cache, True, **kw)
        T._write_back(fmps, k, tn)
        T.sequential_update(fmps, None, cache, k, True)
        d2h += site_bytes(k) + site_bytes(k + 1) + 64
        k += 1
        wall.append(round((time.perf_counter() - w0) * 1e3, 3))
    barrier()
    h2d = eng._h2d_bytes                  # bytes the facade actually uploaded (tensors it handed out itself are not re-sent)
    e2e_value = N * steps / max_over_ranks(time.perf_counter() - e0)
    e2e = {'value': e2e_value, 'unit': 'sample·site updates/s', 'h2d_bytes_per_step': h2d // steps, 'd2h_bytes_per_step': d2h // steps,
           'api': 'facade TNutils.learning_step_cached + write-back + sequential_update, host numpy tensors in and out',
           'host_wall_ms_per_step': wall}

    trace('facade e2e done')
    # ---- e2e (b): the engine-level call with PINNED host site tensors (BornEngine.learning_step_host)
    host = {}
    def pinned(a):
        tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return tt.numpy()
    for j in range(k, k + steps + 1):
        host[j] = pinned(eng.get_site(j))
    h2d = d2h = 0
    barrier()
    e0 = time.perf_counter()
    for i in range(steps):
        a, b = host[k], host[k + 1]
        h2d += a.nbytes + b.nbytes
        na, nb, out = eng.learning_step_host(k, a, b, 1e-3, True, **kw)
        d2h += na.nbytes + nb.nbytes + 64
        host[k + 1] = nb
        k += 1
    barrier()
    rec['e2e_engine_api'] = {'value': N * steps / max_over_ranks(time.perf_counter() - e0), 'unit': 'sample·site updates/s',
                             'h2d_bytes_per_step': h2d // steps, 'd2h_bytes_per_step': d2h // steps,
                             'api': 'BornEngine.learning_step_host, pinned host site tensors'}
    rec['e2e'] = e2e

    trace('engine e2e done')
    # ---- per-kernel pass for the roofline (separate, events around each hot launch)
    eng.set_timing(True)
    for _ in range(max(3, min(steps, 8))):
        eng.step(k, lr, True, **kw); k += 1
    tim = eng.get_timing()
    eng.set_timing(False)
    n_loc = shard.shape[0]
    ker = {}
    for name, flops in (('grad', 2.0 * D * D * n_loc), ('psi', 2.0 * D * D * n_loc), ('env_advance', 2.0 * D * D * n_loc)):   # independent of d
        ms, cnt = tim[name]
        if cnt:
            ker[name] = {'ms': ms / cnt, 'tflops': flops / (ms / cnt * 1e-3) / 1e12}
    for name in ('grad_reduce', 'merge', 'svd'):
        ms, cnt = tim[name]
        if cnt:
            ker[name] = {'ms': ms / cnt}
    dom = max(('grad', 'psi', 'env_advance'), key=lambda n: ker.get(n, {'ms': 0})['ms'])
    peaks, peak_src = load_peaks()
    fp64_peak, fp64_src = fp64_peak_live(f'cuda:{local}')
    traffic, traffic_src = ncu_traffic('k_' + dom, workload if not args.samples else 'custom', world)
    rec['roofline'] = {'bound': 'tensor', 'kernel': 'k_' + dom, 'achieved': ker[dom]['tflops'], 'peak': fp64_peak,
                       'unit': 'TFLOP/s', 'frac': ker[dom]['tflops'] / fp64_peak, 'traffic': traffic, 'traffic_source': traffic_src,
                       'peak_source': f'FP64 {fp64_src} (FP64 is not in MEASURED_PEAKS.json [{peak_src}]; tcgen05 has no FP64 kind: DMMA.8x8x4)',
                       'algorithmic_flops_per_launch': 2.0 * D * D * n_loc,
                       'algorithmic_bytes_per_launch': 2.0 * D * 8 * n_loc, 'kernels': ker,
                       'whole_step_tflops': 6.0 * D * D * n_loc / (dev_ms / steps * 1e-3) / 1e12,
                       'whole_step_frac': 6.0 * D * D * n_loc / (dev_ms / steps * 1e-3) / 1e12 / fp64_peak}

    trace('roofline pass done')
    # ---- a FULL two-site sweep (there and back, 2(L-1) steps over the whole chain incl. the narrow bonds at the ends),
    #      timed on the device as one region: the north star's sweeps/s, measured instead of extrapolated
    if not args.no_sweep:
        mps0 = mpslib.random_mps(L, D, d=d, rng=np.random.default_rng(12), centre=0)
        eng.set_mps(mps0)
        eng.env_build(0)
        e0s, e1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0s.record()
        outs = eng.epoch(1e-3, **kw)
        e1s.record()
        barrier()
        sw_ms = max_over_ranks(e0s.elapsed_time(e1s))
        rec['full_sweep'] = {'seconds': sw_ms * 1e-3, 'steps': len(outs), 'sweeps_per_s': 1e3 / sw_ms,
                             'value': N * len(outs) / (sw_ms * 1e-3), 'unit': 'sample·site updates/s',
                             'skipped_steps': int(sum(1 for o_ in outs if o_['skipped'])),
                             'max_chi': int(max(o_['chi'] for o_ in outs))}

    trace('sweep done')
    # ---- secondary metric of BASELINE.json: samples generated / s (batched ancestral sampler, device-resident).
    #      Reuses this engine: the sampler's arithmetic and cost do not depend on the gauge (config 4 itself is below).
    if world == 1 and not args.no_sampler:
        ns = 20000
        eng.sample_device_ms(2048, seed=3)
        ms_s = [eng.sample_device_ms(ns, seed=7 + i) for i in range(2)]
        rec['sampler'] = {'metric': 'samples_generated_per_s', 'value': ns / (min(ms_s) * 1e-3), 'unit': 'samples/s',
                          'n_samples': ns, 'sites': L, 'Dmax': D, 'ms': min(ms_s),
                          'tflops_executed': 2.0 * D * D * 2.0 * L * ns / (min(ms_s) * 1e-3) / 1e12,
                          'note': 'generate_samples (TNutils.py:1829) on the bench MPS, uniforms generated on device'}
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    rec = measure_steps(args, args.workload, rank, world, local, full=True)
    N, D, d, L = rec['N'], rec['D'], rec['d'], rec['L']

    # ---- the other BASELINE configs that fit one GPU, each as a short run of the same measurement (N=1 only)
    others = {}
    if world == 1 and args.workload == 'c3' and not args.samples and not args.no_others:
        for w in ('c2', 'c5', 'c5b', 'c3steep', 'c3steep_lr1e-9'):
            try:
                r = measure_steps(args, w, rank, world, local, full=False)
                others[w] = {'workload': workload_config(w, 1)['workload'], 'ms_per_step': r['ms_per_step'], 'value': r['value'],
                             'unit': 'sample·site updates/s', 'chi': r['chi'], 'split_eig': r['split_eig'], 'kernel_ms': r['kernel_ms'],
                             'gram_levels': r['gram_levels']}
            except Exception as exc:
                others[w] = {'error': str(exc)}
        try:
            others['c4'] = sampler_record(args, rank, world, local, n_samples=100000, steps=1)
        except Exception as exc:
            others['c4'] = {'error': str(exc)}
        try:
            others['c4_reconstruct'] = reconstruct_record(args, local)
        except Exception as exc:
            others['c4_reconstruct'] = {'error': str(exc)}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cs = max(2, min(8, int(20.0 / max(0.05, 1.5e-5 * N * D / 256))))
            tcpu = cpu_reference_steps(N, D, cs, d=d)
            cpu = {'value': N / float(np.mean(tcpu)), 'unit': 'sample·site updates/s', 'cores': os.cpu_count(), 'kind': 'port',
                   'blas_threads': blas_threads(),
                   'sample': f'{cs} mid-chain two-site steps at N={N}, D={D} (numpy float64 oracle port, BLAS pool sized to all host cores)'}
        line = {
            'metric': 'sample_site_updates_per_s_two_site_sweep', 'value': rec['value'], 'unit': 'sample·site updates/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': rec['ms_per_step'], 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64' if args.precision == 'fp64' else 'f64 (gradient accumulation: tf32 operands, f32 TMEM accumulators)', 'data': 'synthetic',
            'config': workload_config(args.workload, world),
            'e2e': rec['e2e'], 'e2e_engine_api': rec['e2e_engine_api'],
            'gpu_launches': rec['gpu_launches'], 'roofline': rec['roofline'], 'clocks': rec['clocks'],
            'svd': args.svd, 'split_eig': rec['split_eig'], 'allreduce': rec['allreduce'], 'chi': rec['chi'], 'step_ms': rec['step_ms'],
            'gram_levels': rec['gram_levels'], 'wall_s_timed': rec['wall_s_timed'], 'setup_s': rec['setup_s'],
            'sweep_s_extrapolated': rec['ms_per_step'] * 1e-3 * 2 * (L - 1),
        }
        for key in ('full_sweep', 'sampler', 'multi_gpu_parity'):
            if key in rec:
                line[key] = rec[key]
        if cpu:
            line['cpu_baseline'] = cpu
        if others:
            line['other_configs'] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sampler_record(args, rank, world, local, n_samples, steps):
    """Workload c4: batched ancestral sampling (generate_samples, TNutils.py:1829-1872) of N images from a
    right-canonical random MPS with D=500 (no trained D=500 model ships with the reference).  Metric: samples/s.
    One step = one full pass over the 784 sites for all N samples."""
    import torch
    from mps_born_machines_b200 import BornEngine, mpslib
    N, L, D = n_samples, 784, args.sampler_bond
    n_loc = (N + world - 1) // world                      # replicas only: every rank samples its share, no collective
    mps = mpslib.random_mps(L, D, rng=np.random.default_rng(13), centre=0)
    eng = BornEngine(L, D, device=local)
    eng.set_mps(mps)
    for _ in range(max(1, args.warmup)):
        eng.sample_device_ms(min(n_loc, 4096), seed=1)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = eng.launch_count()
    clocks.mark()
    ms = [eng.sample_device_ms(n_loc, seed=100 + i) for i in range(steps)]
    clocks.mark()
    launches = eng.launch_count() - l0
    clk = clocks.stop() if rank == 0 else {}
    t = torch.tensor([sum(ms)], dtype=torch.float64, device=f'cuda:{local}')
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = N * steps / (dev_ms * 1e-3)
    # e2e: host uniforms in (pinned), host samples out, through the public call
    ne = min(n_loc, 16384)
    U = torch.from_numpy(np.random.default_rng(17).random((L, ne))).pin_memory().numpy()
    t0 = time.perf_counter(); out = eng.sample(ne, U=U); e2e_s = time.perf_counter() - t0
    f0 = 1.0 - float(out.mean())
    both = 2.0 * D * D * 2.0 * L * N * steps / (dev_ms * 1e-3) / 1e12         # if both branches were evaluated for every row
    algorithmic = both * (1.0 + f0) / 2.0        # the reference evaluates the second branch only for the rows that reject the first
    executed = both if os.environ.get('BM_SAMPLE_FUSED') == '1' else algorithmic   # two-phase step: second branch for the rejected rows only
    peak, peak_src = fp64_peak_live(f'cuda:{local}')
    eng.close()
    line = None
    if rank == 0:
        from oracle import born_oracle as o
        nc = 2000
        Uc = np.random.default_rng(17).random((8, nc))
        # CPU port of the same per-site work: 8 mid-chain sites at bond D on nc samples, extrapolated to 784 sites
        half = np.random.default_rng(3).standard_normal((nc, D)); m3 = o.site3(mps, 400)
        tc = time.perf_counter()
        for s_ in range(8):
            ampl = half @ m3[:, :, 0]
            cond = np.einsum('nb,nb->n', ampl, ampl) / np.einsum('nb,nb->n', half, half)
            acc = Uc[s_] <= cond
            new = np.where(acc[:, None], ampl, half @ m3[:, :, 1])
            half = new / np.abs(new).max(axis=1, keepdims=True)
        tc = (time.perf_counter() - tc) / 8 * L
        line = {'metric': 'samples_generated_per_s', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': steps,
                'warmup': args.warmup, 'ms_per_step': dev_ms / steps, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': f'c4: batched ancestral sampling of {N} images, L={L}, right-canonical random MPS D={D}, '
                                       'device-generated uniforms (counter-based), second branch only for rejected rows', 'n_samples': N, 'Dmax': D,
                           'parallelism': 'replicas only (no collective)', 'l2': f'state 2x{n_loc}x{D}x8 B ping-pong > L2'},
                'e2e': {'value': ne / e2e_s, 'unit': 'samples/s', 'h2d_bytes_per_step': int(U.nbytes), 'd2h_bytes_per_step': int(out.nbytes)},
                'gpu_launches': int(launches),
                'roofline': {'bound': 'tensor', 'kernel': 'k_sample_step', 'achieved': algorithmic, 'executed': executed, 'peak': peak,
                             'unit': 'TFLOP/s', 'frac': algorithmic / peak, 'frac_executed': executed / peak, 'traffic': None,
                             'peak_source': peak_src,
                             'note': 'achieved = ALGORITHMIC flops 2*D^2*(1+f) per sample and site (f = %.3f: share of rows whose first '
                                     'branch is rejected); the two-phase step computes the second branch for exactly those rows '
                                     '(round 1 evaluated both branches for every 64-row tile holding a rejection)' % f0},
                'clocks': clk,
                'cpu_baseline': {'value': nc / tc, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port', 'blas_threads': blas_threads(),
                                 'sample': f'8 mid-chain sites at D={D} on {nc} samples (numpy float64 port of generate_samples), x{L}/8'}}
    return line


def reconstruct_record(args, local, n_prefix=20000, n_general=64):
    """Config 4, second half: half-image reconstruction (reconstruct, TNutils.py:2053-2116 on
    partial_removal_img(img, (28,28), .5, axis=0, half) masks) at D = --sampler-bond.
      half=0 (bottom half unknown): known prefix -> batched sampler path, n_prefix images in one call;
      half=1 (top half unknown): general mask -> bm_reconstruct_batch (D^3 per site and image: right-environment
      matrices), n_general images in one call.  Host pixels + host uniforms in, host pixels out (e2e by construction)."""
    import torch
    from mps_born_machines_b200 import BornEngine, mpslib
    L, D = 784, args.sampler_bond
    mps = mpslib.random_mps(L, D, rng=np.random.default_rng(13), centre=0)
    eng = BornEngine(L, D, device=local)
    eng.set_mps(mps)
    rng = np.random.default_rng(19)
    out = {}
    imgs = mpslib.synthetic_images(n_prefix, L).astype(np.int64)
    corr = imgs.copy(); corr[:, 392:] = -1
    U = rng.random((392, n_prefix))
    eng.sample(256, U=U[:, :256].copy(), rule='single', prefix=corr[:256], start_site=392)        # warm-up
    torch.cuda.synchronize(); l0 = eng.launch_count(); t0 = time.perf_counter()
    got = eng.sample(n_prefix, U=U, rule='single', prefix=corr, start_site=392)
    dt = time.perf_counter() - t0
    assert np.array_equal(got[:, :392], imgs[:, :392].astype(np.uint8))
    out['bottom_half_unknown'] = {'images': n_prefix, 'seconds': dt, 'images_per_s': n_prefix / dt, 'gpu_launches': eng.launch_count() - l0,
                                  'path': 'known prefix: prefix chain + batched single-rule sampler (bm_sample)'}
    corr = imgs[:n_general].copy(); corr[:, :392] = -1
    U2 = rng.random((n_general, 392))
    eng.reconstruct_batch(corr[:2], U2[:2])                                                        # warm-up
    torch.cuda.synchronize(); l0 = eng.launch_count(); t0 = time.perf_counter()
    got = eng.reconstruct_batch(corr, U2)
    dt = time.perf_counter() - t0
    assert np.array_equal(got[:, 392:], imgs[:n_general, 392:].astype(np.uint8))
    flops = n_general * 391 * 4 * 2.0 * D ** 3
    out['top_half_unknown'] = {'images': n_general, 'seconds': dt, 'images_per_s': n_general / dt, 'gpu_launches': eng.launch_count() - l0,
                               'tflops': flops / dt / 1e12,
                               'path': 'general mask: batched right-environment matrices (bm_reconstruct_batch), 4 D^3 products per unknown site and image'}
    eng.close()
    return out


def run_sampler(args):
    import torch
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    line = sampler_record(args, rank, world, local, args.samples or 100000, args.steps)
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS) + ['c4'])
    ap.add_argument('--samples', type=int, default=0)
    ap.add_argument('--svd', default='gram', choices=['gram', 'gram_syevd', 'gesvdj'])
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-sampler', action='store_true', help='skip the secondary samples/s measurement')
    ap.add_argument('--no-sweep', action='store_true', help='skip the full 2(L-1)-step sweep measurement')
    ap.add_argument('--no-others', action='store_true', help='skip the short runs of BASELINE configs 2, 4, 5')
    ap.add_argument('--allreduce', default='peer', choices=['peer', 'nccl'],
                    help='multi-GPU gradient exchange: fused reduce + peer-memory all-reduce kernel, or NCCL')
    ap.add_argument('--precision', default='fp64', choices=['fp64', 'tf32', 'tf32_reg'],
                    help="fp64 (default, 1e-6 parity) or tf32 (gradient accumulation on tcgen05/TMEM, 1e-3 parity)")
    ap.add_argument('--sampler-bond', type=int, default=500)
    ap.add_argument('--per-step', action='store_true', help='time every step as its own call with an L2 flush in between, also at N=1')
    ap.add_argument('--clock-ms', type=int, default=10, help='nvidia-smi sampling period of the clocks line')
    args = ap.parse_args()
    if args.workload == 'c4':
        run_sampler(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    if os.environ.get('BENCH_TRACE'):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get('BENCH_TRACE_DUMP_S', '90')), exit=True)
    main()
