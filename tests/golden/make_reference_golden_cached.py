"""Golden vectors from the REAL reference's cached (G2) path and batched sampler.

`left_right_cache` (TNutils.py:457), `sequential_update` (:504), `arr_psi_primed_cache` (:1486),
`learning_step_cached` (:1510), `learning_epoch_cached` (:1570), `computeNLL` (:872), `tneinsum3` (:365),
`into_data`/`into_tensarr` (:352-363), `generate_samples` (:1829) are executed UNMODIFIED from /root/reference.
What is replaced is only the absent third-party layer they call into:

  * quimb `Tensor` -> `FakeTensor` below: named-index container; `@`/tensor_contract = np.einsum over shared names,
    arithmetic aligns the other operand's indices by name (as quimb does), `.split` = thin SVD (np.linalg.svd, i.e.
    LAPACK gesdd like quimb's numba path) + THE RESTATED TRUNCATION RULE (cutoff 1e-10 relative, >= 1, max_bond cap,
    no renormalisation, absorb left/right).  The truncation rule therefore stays unpinned; everything around it —
    which cache slot feeds which contraction, index orders, the max-abs rescale, dNLL = A/Z - mean(psi'/psi), the
    update, A/A.data.max(), left_inds/absorb per direction, the write-back transposes, the sweep schedule, the
    `<=` acceptance rule and state update of the batched sampler — is the reference's own code.
  * opt_einsum.contract_expression -> np.einsum on the float32 operands the reference prepares (it asks for
    backend='torch' float32; numpy float32 is the same arithmetic class), get_symbol -> letters.
  * cytoolz.concat, pydash.get, tqdm -> trivial stand-ins.

The reference's caches are float32 (`into_data`), so these fixtures pin the oracle at ~1e-5, not 1e-12.
Run in the build container:  python tests/golden/make_reference_golden_cached.py
"""
import itertools
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
LETTERS = 'abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ'


def _contract(ts):
    sym = {}
    def s(i):
        return sym.setdefault(i, LETTERS[len(sym)])
    ins = [''.join(s(i) for i in t.inds) for t in ts]
    counts = {}
    for t in ts:
        for i in t.inds:
            counts[i] = counts.get(i, 0) + 1
    out_inds = [i for t in ts for i in t.inds if counts[i] == 1]
    out = ''.join(sym[i] for i in out_inds)
    res = np.einsum(','.join(ins) + '->' + out, *[t.data for t in ts])
    return res, tuple(out_inds)


class FakeTensor:
    def __init__(self, data=1.0, inds=(), tags=None):
        self.data = np.asarray(data)
        self.inds = tuple(inds)
        self.tags = tags

    @property
    def shape(self):
        return self.data.shape

    def modify(self, data=None, inds=None, tags=None):
        if data is not None:
            self.data = np.asarray(data)
        if inds is not None:
            self.inds = tuple(inds)

    def _aligned(self, other):
        if isinstance(other, FakeTensor):
            if set(other.inds) != set(self.inds):
                raise ValueError('index mismatch')
            return other.data.transpose([other.inds.index(i) for i in self.inds])
        return other

    def __matmul__(self, other):
        res, inds = _contract([self, other])
        return float(res) if not inds else FakeTensor(res, inds)

    def __truediv__(self, x): return FakeTensor(self.data / self._aligned(x), self.inds)
    def __mul__(self, x): return FakeTensor(self.data * self._aligned(x), self.inds)
    __rmul__ = __mul__
    def __sub__(self, x): return FakeTensor(self.data - self._aligned(x), self.inds)
    def __add__(self, x): return FakeTensor(self.data + self._aligned(x), self.inds)
    __radd__ = __add__

    def split(self, left_inds, absorb='both', max_bond=None, cutoff=1e-10, **kw):
        left_inds = list(left_inds)
        right_inds = [i for i in self.inds if i not in left_inds]
        x = self.data.transpose([self.inds.index(i) for i in left_inds + right_inds])
        lshape, rshape = x.shape[:len(left_inds)], x.shape[len(left_inds):]
        U, s, Vh = np.linalg.svd(x.reshape(int(np.prod(lshape)), int(np.prod(rshape))), full_matrices=False)
        # ---- restated quimb 1.3.0 rule (cutoff_mode='rel', renorm off) — the unpinned part
        if cutoff > 0.0:
            chi = max(int(np.sum(s > cutoff * s[0])), 1)
            if max_bond is not None and max_bond > 0:
                chi = min(chi, max_bond)
        elif max_bond is not None and max_bond > 0:
            chi = min(max_bond, s.size)
        else:
            chi = s.size
        U, s, Vh = U[:, :chi], s[:chi], Vh[:chi]
        if absorb == 'right':
            Vh = s[:, None] * Vh
        elif absorb == 'left':
            U = U * s[None, :]
        else:
            raise NotImplementedError(absorb)
        tn = types.SimpleNamespace()
        tn.tensors = (FakeTensor(U.reshape(lshape + (chi,)), tuple(left_inds) + ('_bond',)),
                      FakeTensor(Vh.reshape((chi,) + rshape), ('_bond',) + tuple(right_inds)))
        return tn


class FakeMPS:
    def __init__(self, arrays):
        L = len(arrays)
        self.tensors = []
        for k, a in enumerate(arrays):
            inds = ('i0', 'v0') if k == 0 else ((f'i{k - 1}', f'v{k}') if k == L - 1 else (f'i{k - 1}', f'i{k}', f'v{k}'))
            self.tensors.append(FakeTensor(np.array(a, dtype=np.float64), inds))

    def __getitem__(self, k):
        return self.tensors[k]

    def __matmul__(self, other):
        raise NotImplementedError


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Any()
        def __getattr__(self, k): return _Any()

    def gen_output_inds(all_inds):
        all_inds = list(all_inds)
        counts = {}
        for i in all_inds:
            counts[i] = counts.get(i, 0) + 1
        seen = set()
        for i in all_inds:
            if counts[i] == 1 and i not in seen:
                seen.add(i)
                yield i

    def contract_expression(eq, *shapes):
        def expr(*arrays, backend=None):
            return np.einsum(eq, *arrays)          # operands are the float32 arrays the reference built
        return expr

    class _Progress(list):
        def set_description(self, *_a, **_k): pass

    core = types.SimpleNamespace(_gen_output_inds=gen_output_inds)
    _stub('cytoolz', concat=itertools.chain.from_iterable)
    _stub('dask', array=None)
    _stub('dask.array', Array=type('Array', (), {}))
    sys.modules['dask'].array = sys.modules['dask.array']
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.pylab']:
        _stub(name)
    sys.modules['matplotlib.colors'].LinearSegmentedColormap = _Any
    sys.modules['matplotlib.colors'].ListedColormap = _Any
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['matplotlib'].pylab = sys.modules['matplotlib.pylab']
    qt = _stub('quimb.tensor', Tensor=FakeTensor, tensor_core=core,
               tensor_contract=lambda *ts: (lambda r, i: float(r) if not i else FakeTensor(r, i))(*_contract(list(ts))))
    _stub('quimb', tensor=qt)
    _stub('opt_einsum', contract_expression=contract_expression, get_symbol=lambda i: LETTERS[i])
    _stub('pydash', get=lambda obj, key, default=None: getattr(obj, key, default))
    _stub('tqdm.notebook', tqdm=lambda it, **k: _Progress(it))
    try:
        import torchvision  # noqa
    except Exception:
        _stub('torchvision'); _stub('torchvision.transforms')
    sys.path.insert(0, REF)
    import TNutils
    return TNutils


def right_canonical_mps(L, D, seed):
    rng = np.random.default_rng(seed)
    mats = [rng.standard_normal((1 if k == 0 else D, 1 if k == L - 1 else D, 2)) for k in range(L)]
    for k in range(L - 1, 0, -1):
        dl, dr, d = mats[k].shape
        q, r = np.linalg.qr(mats[k].reshape(dl, dr * d).T)
        mats[k] = q.T.reshape(-1, dr, d)
        mats[k - 1] = np.einsum('abp,bc->acp', mats[k - 1], r.T)
    mats[0] = mats[0] / np.linalg.norm(mats[0])
    return [np.ascontiguousarray(t[0]) if k == 0 else (np.ascontiguousarray(t[:, 0, :]) if k == L - 1 else np.ascontiguousarray(t))
            for k, t in enumerate(mats)]


def main():
    import contextlib, io
    T = import_reference()
    L, D, N = 8, 3, 14
    arrays = right_canonical_mps(L, D, 77)
    rng = np.random.default_rng(78)
    imgs = (rng.random((N, L)) < 0.45).astype(np.int64)
    fix = {'L': L, 'imgs': imgs}
    for k, a in enumerate(arrays):
        fix[f'site{k}'] = a

    # ---- left_right_cache (float32, max-abs rescaled)
    mps = FakeMPS(arrays)
    _imgs = np.array([T.tens_picture(img) for img in imgs])
    cache = T.left_right_cache(mps, _imgs)
    for j in range(L):
        fix[f'cacheL_{j}'] = np.stack([np.atleast_1d(np.asarray(cache[n, 0, j].data, dtype=np.float64)) for n in range(N)])
        fix[f'cacheR_{j}'] = np.stack([np.atleast_1d(np.asarray(cache[n, 1, j].data, dtype=np.float64)) for n in range(N)])

    # ---- one learning_step_cached at bond 3, going right (teacher-forced quantities)
    SD = T.learning_step_cached(mps, 3, _imgs, 0.05, cache, True, max_bond=5)
    fix['step3_left'] = SD.tensors[0].data
    fix['step3_right'] = SD.tensors[1].data
    fix['step3_prod'] = np.einsum('apb,bcq->apcq', SD.tensors[0].data, SD.tensors[1].data)

    # ---- learning_epoch_cached: 3 epochs, full batch, max_bond 6 (reference signature with val_imgs)
    mps = FakeMPS(arrays)
    cache = T.left_right_cache(mps, _imgs)
    np.random.seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        cost, lr = T.learning_epoch_cached(mps, imgs, _imgs, 3, 0.1, cache, batch_size=N, lr_update=lambda x: x * 0.9, max_bond=6)
    fix['epoch_cost'] = np.array(cost, dtype=np.float64)
    fix['epoch_lr'] = lr
    fix['epoch_bonds'] = np.array([mps.tensors[k].data.shape[1 if k else 0] for k in range(L - 1)])
    for k in range(L):
        fix[f'trained{k}'] = np.asarray(mps.tensors[k].data, dtype=np.float64)

    # ---- batched sampler on the trained (right-canonical, centre 0) MPS; float32 amplitudes in the reference
    Ns = 300
    np.random.seed(11)
    samples = T.generate_samples(mps, Ns)
    np.random.seed(11)
    U = np.stack([np.random.random(Ns) for _ in range(L)])
    fix['samples'] = samples
    fix['uniforms'] = U
    np.savez_compressed(os.path.join(HERE, 'ref_cached_path.npz'), **fix)
    print('wrote ref_cached_path.npz; epoch costs', cost, 'bonds', fix['epoch_bonds'])


if __name__ == '__main__':
    main()
