"""BornEngine — Python host wrapper over libbm.so (include/bm.h).

One engine per process/GPU.  All arithmetic of the hot path runs in the CUDA library; PyTorch is used for
the gradient buffer (so that torch.distributed can all-reduce it over NCCL) and for stream plumbing.
Reference functions replaced (TNutils.py): left_right_cache :457, sequential_update :504,
learning_step_cached :1510, learning_epoch_cached :1570, computeNLL :872, generate_samples :1829,
generate_sample :1874 / reconstruct :2053 (known-prefix case).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _capi
from .parallel import allreduce_sum_
from ._capi import RENORM, RULE, SVD, check, lib

SCALAR_NAMES = ('Z', 'sum_logpsi', 'n_zero_psi', 'chi', 'mean_dnll', 'renorm_div', 'trunc_err', 'nonfinite')


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def sweep_schedule(L):
    """[0..L-2] + [L-2..0]  (TNutils.py:1597)."""
    return list(range(0, L - 1)) + list(range(L - 2, -1, -1))


class BornEngine:
    _advanced = None      # (bond, going_right) of the last step whose environment advance is already enqueued (facade)

    def __init__(self, n_sites, max_bond, phys_dim=2, device=None, svd='gram', dist_group=None, precision='fp64',
                 standalone=False):
        if not torch.cuda.is_available():
            raise RuntimeError('BornEngine needs a CUDA device (sm_100a); there is no CPU fallback')
        self.L, self.d, self.Dcap = int(n_sites), int(phys_dim), int(max_bond)
        self.device = torch.cuda.current_device() if device is None else int(device)
        torch.cuda.set_device(self.device)
        self._lib = lib()
        self._ctx = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        code = self._lib.bm_create(C.byref(self._ctx), self.device, self.L, self.d, self.Dcap, C.c_void_p(stream))
        check(self._ctx, code)
        check(self._ctx, self._lib.bm_set_svd(self._ctx, SVD[svd]))
        if precision not in ('fp64', 'tf32', 'tf32_reg'):
            raise ValueError("precision must be 'fp64', 'tf32' (TMA-staged tcgen05 gradient) or 'tf32_reg' (register-staged)")
        check(self._ctx, self._lib.bm_set_precision(self._ctx, {'fp64': 0, 'tf32_reg': 1, 'tf32': 2}[precision]))
        self.precision = precision
        ld = (self.Dcap + 1) & ~1
        self._S = torch.zeros(self.d * self.Dcap * self.d * ld + 2, dtype=torch.float64, device=f'cuda:{self.device}')
        self.N = 0
        self.N_total = 0
        self.peer = False
        self.group = dist_group
        self.world = 1
        # standalone=True: an engine that holds ALL its data and never takes part in a collective, even inside a
        # multi-rank job (the single-GPU reference of the multi-GPU parity checks)
        if not standalone and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(dist_group)

    def enable_peer_allreduce(self):
        """Switch the per-step gradient all-reduce from NCCL to the fused peer-memory kernel (k_grad_reduce_peer):
        exchanges CUDA IPC handles of the per-rank exchange buffers once; afterwards step_grad returns the
        already all-reduced buffer and step() issues no collective."""
        if self.world <= 1:
            return False
        import torch.distributed as dist
        ok, err = 1, None
        try:
            h = (C.c_char * 64)()
            check(self._ctx, self._lib.bm_peer_export(self._ctx, h))
            payload = bytes(h.raw)
        except Exception as exc:      # keep going: the outcome must be agreed on by ALL ranks
            ok, err, payload = 0, exc, b'\0' * 64
        gathered = [None] * self.world
        dist.all_gather_object(gathered, payload, group=self.group)
        rank = dist.get_rank(self.group)
        if ok:
            try:
                blob = b''.join(gathered)
                buf = (C.c_char * len(blob)).from_buffer_copy(blob)
                check(self._ctx, self._lib.bm_peer_import(self._ctx, buf, self.world, rank))
            except Exception as exc:
                ok, err = 0, exc
        flag = torch.tensor([ok], dtype=torch.int32, device=f'cuda:{self.device}')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:     # some rank could not map its peers: everybody stays on the NCCL path
            one = (C.c_char * 64)()
            self._lib.bm_peer_import(self._ctx, one, 1, 0)
            self.peer = False
            raise RuntimeError(f'peer-memory all-reduce unavailable on at least one rank ({err})')
        self.peer = True
        return True

    def close(self):
        if getattr(self, '_ctx', None) is not None and self._ctx.value:
            self._lib.bm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data / model ----------------------------------------------------------------------------
    def set_data(self, imgs, n_total=None):
        """imgs: (N,L) pixel levels 0..d-1 of THIS rank's shard.  n_total: global sample count."""
        a = np.ascontiguousarray(np.asarray(imgs), dtype=np.int64)
        if a.ndim != 2 or a.shape[1] != self.L:
            raise ValueError(f'expected (N,{self.L}) images')
        if a.min() < 0 or a.max() >= self.d:
            raise ValueError('training images must hold levels 0..d-1 (no unknown pixels)')
        px = np.ascontiguousarray(a.astype(np.uint8))
        # every rank learns every shard size: the global batch size of a step must be the same number on all ranks
        # (shards may be uneven), otherwise the replicated update diverges silently
        self.shard_sizes = [px.shape[0]]
        if self.world > 1:
            sizes = [None] * self.world
            torch.distributed.all_gather_object(sizes, int(px.shape[0]), group=self.group)
            self.shard_sizes = [int(x) for x in sizes]
        if min(self.shard_sizes) < 1:
            raise ValueError(f'every rank needs at least one training sample (shard sizes {self.shard_sizes}); '
                             'use fewer ranks or parallel.shard_bounds on a larger data set')
        check(self._ctx, self._lib.bm_set_data(self._ctx, _ptr(px), px.shape[0]))
        self.N = px.shape[0]
        self.N_total = sum(self.shard_sizes)
        if n_total is not None and int(n_total) != self.N_total:
            raise ValueError(f'n_total={n_total} does not match the sum of the shard sizes {self.shard_sizes}')

    def set_site(self, k, t):
        self._advanced = None
        t = np.asarray(t, dtype=np.float64)
        if t.ndim == 2:
            t = t.reshape(1, t.shape[0], t.shape[1]) if k == 0 else t.reshape(t.shape[0], 1, t.shape[1])
        t = np.ascontiguousarray(t)
        if t.shape[2] != self.d:
            raise ValueError('physical dimension mismatch')
        # NaN and inf always survive a sum; only a non-finite sum pays for the exact elementwise test
        if not math.isfinite(float(t.sum())) and not np.isfinite(t).all():
            raise ValueError(f'site tensor {k} holds non-finite entries')
        check(self._ctx, self._lib.bm_set_site(self._ctx, k, _ptr(t), t.shape[0], t.shape[1]))

    def set_mps(self, mps):
        if len(mps) != self.L:
            raise ValueError('wrong number of sites')
        for k, t in enumerate(mps):
            self.set_site(k, t)

    def site_dims(self, k):
        dl, dr = C.c_int(), C.c_int()
        check(self._ctx, self._lib.bm_site_dims(self._ctx, k, C.byref(dl), C.byref(dr)))
        return dl.value, dr.value

    def get_site(self, k, squeeze=True):
        dl, dr = self.site_dims(k)
        out = np.empty((dl, dr, self.d))
        check(self._ctx, self._lib.bm_get_site(self._ctx, k, _ptr(out), None, None))
        if squeeze and k == 0:
            return out[0]
        if squeeze and k == self.L - 1:
            return out[:, 0, :]
        return out

    def get_mps(self):
        return [self.get_site(k) for k in range(self.L)]

    def bond_sizes(self):
        return [self.site_dims(k)[1] for k in range(self.L - 1)]

    # ---- environments ----------------------------------------------------------------------------
    def env_build(self, k=0):
        check(self._ctx, self._lib.bm_env_build(self._ctx, k))

    def env_advance(self, k, going_right):
        check(self._ctx, self._lib.bm_env_advance(self._ctx, k, 1 if going_right else 0))

    def env_get(self, side, j):
        D = self.site_dims(j)[0] if 0 < j < self.L else 1
        out = np.empty((self.N, D))
        got = C.c_int()
        check(self._ctx, self._lib.bm_env_get(self._ctx, side, j, _ptr(out), C.byref(got)))
        assert got.value == D
        return out

    # ---- two-site step ---------------------------------------------------------------------------
    def grad_dims(self, k):
        v = [C.c_int() for _ in range(5)]
        check(self._ctx, self._lib.bm_grad_dims(self._ctx, k, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)   # rows, ld, Dl, Dr, ldR

    def step_grad(self, k, mask=None):
        """merge + psi + S = sum psi'/psi on this rank's samples; returns the device buffer view
        (rows*ld + 2 doubles: S, sum ln|psi|, n_zero)."""
        if mask is not None:
            m = np.ascontiguousarray(np.asarray(mask), dtype=np.int32)
            code = self._lib.bm_step_grad(self._ctx, k, _ptr(m), m.size, C.c_void_p(self._S.data_ptr()))
        else:
            code = self._lib.bm_step_grad(self._ctx, k, None, 0, C.c_void_p(self._S.data_ptr()))
        check(self._ctx, code)
        rows, ld = self.grad_dims(k)[:2]
        return self._S[:rows * ld + 2]

    def step_update(self, k, S, batch_total, lr, going_right=True, renorm='max', max_bond=None, cutoff=1e-10,
                    mom_bias=0.0, want_s=False):
        self._advanced = None
        rows, ld, Dl, Dr, _ = self.grad_dims(k)
        scal = np.zeros(8)
        s = np.empty(min(rows, self.d * Dr)) if want_s else None
        code = self._lib.bm_step_update(self._ctx, k, C.c_void_p(S.data_ptr()), float(batch_total), float(lr),
                                        RENORM[renorm], 1 if going_right else 0,
                                        int(max_bond) if max_bond else 0, float(cutoff), float(mom_bias),
                                        _ptr(scal), _ptr(s) if want_s else None)
        check(self._ctx, code)
        return self._step_result(scal, s, want_s)

    @staticmethod
    def _step_result(scal, s, want_s):
        out = dict(zip(SCALAR_NAMES, scal))
        out['chi'] = int(out['chi'])
        if want_s:
            out['s_all'] = s
            out['s'] = s[:out['chi']]
        # psi == 0 for some sample: the descent was skipped (lr = 0) but the tensor was renormalised and split,
        # exactly as the reference does (TNutils.py:1244-1254)
        out['skipped'] = out['n_zero_psi'] > 0
        return out

    def step(self, k, lr, going_right=True, renorm='max', max_bond=None, cutoff=1e-10, mask=None, mom_bias=0.0,
             want_s=False, batch_total=None, advance=True, stage_sites=False):
        """One learning_step_cached + write-back + sequential_update (TNutils.py:1603-1620).  Single GPU and the
        fused peer all-reduce run the whole step inside ONE library call (bm_step); the NCCL path needs the collective
        between the gradient and the update and goes through step_grad / step_update.  The call returns while the
        environment advance is still running on the stream; `stage_sites` leaves the two new tensors ready for
        `get_site` (a plain copy that overlaps the advance)."""
        self._advanced = None
        if batch_total is None:
            if mask is not None and self.world > 1:
                raise ValueError('multi-rank minibatch steps need batch_total (the global number of samples used by '
                                 'this step, identical on every rank): see BornEngine.epoch')
            batch_total = self.N_total if mask is None else len(mask)
        if self.world > 1 and not self.peer:
            S = self.step_grad(k, mask)
            allreduce_sum_(S, self.group)   # sum over sample shards (NCCL over NVLink)
            out = self.step_update(k, S, batch_total, lr, going_right, renorm, max_bond, cutoff, mom_bias, want_s)
            if advance:
                self.env_advance(k, going_right)
        else:
            rows, ld, Dl, Dr, _ = self.grad_dims(k)
            scal = np.zeros(8)
            s = np.empty(min(rows, self.d * Dr)) if want_s else None
            m = None
            if mask is not None:
                m = np.ascontiguousarray(np.asarray(mask), dtype=np.int32)
            code = self._lib.bm_step(self._ctx, k, _ptr(m) if m is not None else None, m.size if m is not None else 0,
                                     C.c_void_p(self._S.data_ptr()), float(batch_total), float(lr), RENORM[renorm],
                                     1 if going_right else 0, int(max_bond) if max_bond else 0, float(cutoff),
                                     float(mom_bias), (1 if advance else 0) | (2 if stage_sites else 0), _ptr(scal),
                                     _ptr(s) if want_s else None)
            check(self._ctx, code)
            out = self._step_result(scal, s, want_s)
        out['nll_batch'] = -2.0 * out['sum_logpsi'] / batch_total + math.log(out['Z']) if out['Z'] > 0 else float('nan')
        return out

    def split_bond(self, k, going_right=False, max_bond=None, cutoff=1e-10, want_s=False):
        """merge + truncated split of bond k, no gradient step (compress, TNutils.py:930-947)."""
        self._advanced = None
        rows, ld, Dl, Dr, _ = self.grad_dims(k)
        scal = np.zeros(8)
        s = np.empty(min(rows, self.d * Dr)) if want_s else None
        check(self._ctx, self._lib.bm_split_bond(self._ctx, k, 1 if going_right else 0, int(max_bond) if max_bond else 0,
                                                 float(cutoff), _ptr(scal), _ptr(s) if want_s else None))
        out = dict(zip(SCALAR_NAMES, scal))
        out['chi'] = int(out['chi'])
        if want_s:
            out['s'] = s[:out['chi']]
        return out

    def canonize(self, centre):
        """Move the orthogonality centre to `centre` on the device (quimb canonize, TNutils.py:398) by exact
        (untruncated) splits: sites < centre become left isometries, sites > centre right isometries."""
        for k in range(0, centre):
            self.split_bond(k, going_right=True, max_bond=self.site_dims(k)[1], cutoff=0.0)
        for k in range(self.L - 2, centre - 1, -1):
            self.split_bond(k, going_right=False, max_bond=self.site_dims(k)[1], cutoff=0.0)

    def learning_step_host(self, k, site_k, site_k1, lr, going_right=True, **kw):
        """Reference-style call with HOST tensors in and out (learning_step_cached + write-back +
        sequential_update, TNutils.py:1603-1620): uploads the two site tensors, runs the step on the
        device-resident environments, downloads the two new site tensors."""
        self.set_site(k, site_k)
        self.set_site(k + 1, site_k1)
        out = self.step(k, lr, going_right, stage_sites=True, **kw)
        return self.get_site(k), self.get_site(k + 1), out

    def step_with_update_fn(self, k, lr, fn, going_right=True, renorm='max', max_bond=None, cutoff=1e-10, mask=None,
                            advance=True):
        """Two-site step with an arbitrary host-side `update_wrap(site, dNLL) -> update` (TNutils.py:1541).
        dNLL = A/Z - S/B is formed on the host from the device results, `fn` is applied, and an equivalent
        S' = B (A/Z - update) is uploaded so that the device update A - lr*(A/Z - S'/B) equals A - lr*update."""
        self._advanced = None
        S = self.step_grad(k, mask)
        if self.world > 1 and not self.peer:
            allreduce_sum_(S, self.group)
        if mask is not None and self.world > 1:
            raise ValueError('step_with_update_fn: multi-rank minibatches are not supported (use BornEngine.epoch)')
        B = self.N_total if mask is None else len(mask)
        A = self.get_merged_at(k)
        Sref = self.grad_to_ref(k, S)
        Z = float(np.sum(A * A))
        upd = np.asarray(fn(k, A / Z - Sref / B), dtype=np.float64)
        S2 = np.ascontiguousarray(B * (A / Z - upd))
        check(self._ctx, self._lib.bm_ref_to_engine(self._ctx, k, _ptr(S2), C.c_void_p(S.data_ptr())))
        out = self.step_update(k, S, B, lr, going_right, renorm, max_bond, cutoff)
        if advance:
            self.env_advance(k, going_right)
        return out

    def get_merged(self):
        """merged tensor A of the last step_grad, reference index order (Dl,d,Dr,d) (TNutils.py:1388)."""
        raise RuntimeError('use get_merged_at(k) right after step_grad(k)')

    def get_merged_at(self, k):
        rows, ld, Dl, Dr, _ = self.grad_dims(k)
        out = np.empty((Dl, self.d, Dr, self.d))
        check(self._ctx, self._lib.bm_get_merged(self._ctx, _ptr(out)))
        return out

    def grad_to_ref(self, k, S):
        """engine-layout device buffer -> (Dl,d,Dr,d) numpy array."""
        rows, ld, Dl, Dr, _ = self.grad_dims(k)
        out = np.empty((Dl, self.d, Dr, self.d))
        check(self._ctx, self._lib.bm_engine_to_ref(self._ctx, k, C.c_void_p(S.data_ptr()), _ptr(out)))
        return out

    def sweep(self, bonds, going_right, lr, renorm='max', max_bond=None, cutoff=1e-10, masks=None, mom_bias=None,
              batch_total=None):
        """A run of two-site steps in ONE library call (bm_sweep): bonds[i] with direction going_right[i] (a bool or a
        sequence), full batch or one mask per step (masks: (n_steps, n_mask) local sample ids).  Single GPU or the fused
        peer all-reduce only (the torch.distributed all-reduce sits between the gradient and the update of every step).
        Returns the list of per-step result dicts of `step` (without singular values)."""
        self._advanced = None
        if self.world > 1 and not self.peer:
            raise RuntimeError('sweep() needs the fused peer all-reduce with several ranks; use step()/epoch()')
        b = np.ascontiguousarray(np.asarray(bonds, dtype=np.int32))
        n = b.size
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(going_right, dtype=np.int32), (n,)))
        m = None
        if masks is not None:
            m = np.ascontiguousarray(np.asarray(masks, dtype=np.int32).reshape(n, -1))
            if self.world > 1 and batch_total is None:
                raise ValueError('multi-rank minibatch sweeps need batch_total')
        if batch_total is None:
            batch_total = self.N_total if m is None else m.shape[1]
        mom = None if mom_bias is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mom_bias, dtype=np.float64), (n,)))
        scal = np.zeros((n, 8))
        done = C.c_int(0)
        code = self._lib.bm_sweep(self._ctx, _ptr(b), _ptr(g), n, _ptr(m) if m is not None else None, m.shape[1] if m is not None else 0,
                                  C.c_void_p(self._S.data_ptr()), float(batch_total), float(lr), RENORM[renorm],
                                  int(max_bond) if max_bond else 0, float(cutoff), _ptr(mom) if mom is not None else None,
                                  _ptr(scal), C.byref(done))
        check(self._ctx, code)
        outs = []
        for i in range(n):
            o = self._step_result(scal[i], None, False)
            o['nll_batch'] = -2.0 * o['sum_logpsi'] / batch_total + math.log(o['Z']) if o['Z'] > 0 else float('nan')
            outs.append(o)
        return outs

    def epoch(self, lr, renorm='max', max_bond=None, cutoff=1e-10, batch_size=None, rng=None, callback=None):
        """One reference epoch: 2(L-1) steps over the schedule, centre ends at site 0 (TNutils.py:1592-1625).  Without a
        per-step callback the whole epoch is ONE library call (bm_sweep)."""
        L = self.L
        if callback is None and (self.world == 1 or self.peer):
            sched = sweep_schedule(L)
            dirs = [1] * (L - 1) + [0] * (L - 1)
            use_mask = batch_size is not None and any(batch_size < n for n in self.shard_sizes)
            masks, batch_total = None, None
            if use_mask:
                batch_total = sum(min(batch_size, n) for n in self.shard_sizes)
                if batch_size < self.N:
                    gen = rng if rng is not None else np.random
                    masks = np.stack([gen.permutation(self.N)[:batch_size] for _ in sched])
            return self.sweep(sched, dirs, lr, renorm, max_bond, cutoff, masks, None, batch_total)
        going_right = True
        outs = []
        # minibatches: `batch_size` samples PER RANK (all of a shard that is smaller).  Whether masks are used and the
        # global batch size are derived from ALL shard sizes, so every rank applies the same 1/B to the summed gradient.
        use_mask = batch_size is not None and any(batch_size < n for n in self.shard_sizes)
        batch_total = sum(min(batch_size, n) for n in self.shard_sizes) if use_mask else None
        for k in sweep_schedule(L):
            mask = None
            if use_mask and batch_size < self.N:
                guide = (rng if rng is not None else np.random).permutation(self.N)
                mask = guide[:batch_size]
            o = self.step(k, lr, going_right, renorm, max_bond, cutoff, mask, batch_total=batch_total)
            if callback is not None:
                callback(k, going_right, o)
            outs.append(o)
            if k == L - 2:
                going_right = False
        return outs

    # ---- likelihood ------------------------------------------------------------------------------
    def logpsi(self, imgs):
        a = np.ascontiguousarray(np.asarray(imgs), dtype=np.int64)
        px = np.where(a < 0, 255, a).astype(np.uint8)
        px = np.ascontiguousarray(px)
        ln = np.empty(px.shape[0]); sg = np.empty(px.shape[0])
        check(self._ctx, self._lib.bm_logpsi(self._ctx, _ptr(px), px.shape[0], _ptr(ln), _ptr(sg)))
        return sg, ln

    def last_logpsi(self):
        out = np.empty(self.N)
        check(self._ctx, self._lib.bm_last_logpsi(self._ctx, _ptr(out)))
        return out

    def norm2_centre(self, k=0):
        t = self.get_site(k)
        return float(np.sum(t * t))

    def nll(self, imgs, centre=0):
        """computeNLL(mps, imgs, canonicalized_index) (TNutils.py:872-890)."""
        _, ln = self.logpsi(imgs)
        return float(-2.0 * ln.sum() / len(ln) + math.log(self.norm2_centre(centre)))

    # ---- sampling --------------------------------------------------------------------------------
    def sample(self, n, U=None, seed=0, rule='batched', prefix=None, start_site=0):
        """generate_samples (rule='batched', TNutils.py:1829) or the generate_sample/reconstruct rule
        (rule='single', :1874) for n samples in lock step.  U: (L-start_site, n) uniforms, site-major."""
        out = np.empty((n, self.L), dtype=np.uint8)
        up = None
        if U is not None:
            U = np.ascontiguousarray(np.asarray(U, dtype=np.float64))
            if U.shape != (self.L - start_site, n):
                raise ValueError(f'U must have shape ({self.L - start_site},{n})')
            up = _ptr(U)
        pp = None
        if prefix is not None:
            a = np.asarray(prefix)
            px = np.ascontiguousarray(np.where(a < 0, 255, a).astype(np.uint8))
            if px.shape != (n, self.L):
                raise ValueError('prefix images must be (n,L)')
            pp = _ptr(px)
        check(self._ctx, self._lib.bm_sample(self._ctx, n, RULE[rule], start_site, up, C.c_uint64(seed), pp, _ptr(out)))
        return out

    def reconstruct_one(self, img, u):
        """reconstruct (TNutils.py:2053-2116) for one image with -1 (unknown) pixels anywhere; `u`: one uniform per
        unknown pixel in site order.  Returns the completed image."""
        a = np.asarray(img)
        px = np.ascontiguousarray(np.where(a < 0, 255, a).astype(np.uint8))
        uu = np.ascontiguousarray(np.asarray(u, dtype=np.float64))
        out = np.empty(self.L, dtype=np.uint8)
        check(self._ctx, self._lib.bm_reconstruct_one(self._ctx, _ptr(px), _ptr(uu) if uu.size else None, int(uu.size), _ptr(out)))
        return out

    def reconstruct_batch(self, imgs, U, max_images=None):
        """`reconstruct` (TNutils.py:2053-2116) for a batch of images sharing one mask: imgs (B,L) with -1/255 for unknown
        pixels, U (B, n_unknown) uniforms (image b's draws in site order).  Returns (B,L) uint8.  The batch is split so
        that the right-environment matrices of a chunk (chunk * n_unknown * Dcap^2 * 8 bytes) fit in free memory."""
        a = np.asarray(imgs)
        if a.ndim != 2 or a.shape[1] != self.L:
            raise ValueError(f'expected (B,{self.L}) images')
        px = np.ascontiguousarray(np.where(a < 0, 255, a).astype(np.uint8))
        B = px.shape[0]
        unk = px[0] == 255
        if not np.array_equal(px == 255, np.broadcast_to(unk, px.shape)):
            raise ValueError('all images of a batch must share one mask (group them by mask first)')
        nu = int(unk.sum())
        uu = np.ascontiguousarray(np.asarray(U, dtype=np.float64).reshape(B, nu))
        out = np.empty((B, self.L), dtype=np.uint8)
        if max_images is None:
            free, _ = torch.cuda.mem_get_info(self.device)
            ld = (self.Dcap + 1) & ~1
            per = (nu + 3) * self.Dcap * ld * 8
            max_images = int(max(1, min(B, (free * 0.8) // max(per, 1))))
        for lo in range(0, B, max_images):
            hi = min(B, lo + max_images)
            check(self._ctx, self._lib.bm_reconstruct_batch(self._ctx, _ptr(px[lo:hi]), hi - lo, _ptr(uu[lo:hi]) if nu else None,
                                                            nu, _ptr(out[lo:hi])))
        return out

    def sample_device_ms(self, n, seed=0, rule='batched'):
        ms = C.c_float()
        check(self._ctx, self._lib.bm_sample_device(self._ctx, n, RULE[rule], C.c_uint64(seed), C.byref(ms)))
        return ms.value

    TIMING_SLOTS = ('env_advance', 'psi', 'grad', 'grad_reduce', 'merge', 'svd', 'unused', 'sample_step')

    def set_timing(self, on):
        check(self._ctx, self._lib.bm_set_timing(self._ctx, 1 if on else 0))

    def get_timing(self):
        ms = (C.c_double * 8)(); n = (C.c_longlong * 8)()
        check(self._ctx, self._lib.bm_get_timing(self._ctx, ms, n))
        return {name: (ms[i], n[i]) for i, name in enumerate(self.TIMING_SLOTS)}

    def launch_count(self):
        return int(self._lib.bm_launch_count(self._ctx))

    def eig_sym(self, G, m=None, on_chip=True):
        """Eigen-decomposition of a symmetric matrix with the split's eigensolver (`bm_eig_sym`).
        Returns (lam ascending, U = eigenvectors of the m largest in descending order, info dict)."""
        G = np.asfortranarray(G, dtype=np.float64)
        n = G.shape[0]
        m = n if m is None else int(m)
        lam = np.empty(n)
        U = np.empty((n, m), order='F')
        info = np.zeros(8)
        check(self._ctx, self._lib.bm_eig_sym(self._ctx, n, G.ctypes.data, m, int(bool(on_chip)), lam.ctypes.data,
                                              U.ctypes.data, info.ctypes.data))
        return lam, U, {'ms': info[0], 'ortho_dev': info[1], 'on_chip': bool(info[2]),
                        'stage_ms': dict(zip(('sytrd', 'eigvals', 'vectors', 'reflectors', 'polish'), info[3:8].tolist()))}

    FALLBACK_REASONS = ('cluster_unavailable', 'nonfinite', 'zero_matrix', 'too_many_levels', 'unresolved_kept_values',
                        'degenerate_kept_eigenvalues', 'orthonormality_check_failed', 'unused')

    def eig_fallback_reasons(self):
        """why splits left the on-chip eigensolver (counts since the engine was created)."""
        out = (C.c_longlong * 8)()
        check(self._ctx, self._lib.bm_eig_fallback_reasons(self._ctx, out))
        return {k: int(v) for k, v in zip(self.FALLBACK_REASONS, out) if v}

    def eig_stats(self):
        """(splits that ran on the on-chip eigensolver, splits that fell back to cuSOLVER syevd)"""
        a, b, lv = C.c_longlong(), C.c_longlong(), C.c_int()
        check(self._ctx, self._lib.bm_eig_stats(self._ctx, C.byref(a), C.byref(b), C.byref(lv)))
        self.last_gram_levels = lv.value
        return a.value, b.value


def uniforms_reference(seed, site, n):
    """Host mirror of the device generator k_fill_uniform (splitmix64 counter hash)."""
    M = (1 << 64) - 1

    def sm(x):
        x = (x + 0x9E3779B97F4A7C15) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)
    base = sm(seed ^ ((0xD1B54A32D192ED03 * (site + 1)) & M))
    return np.array([(sm((base + i) & M) >> 11) / 9007199254740992.0 for i in range(n)])
