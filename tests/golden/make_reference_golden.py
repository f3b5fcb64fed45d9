"""Generate golden vectors by executing the REAL reference (`/root/reference/TNutils.py`).

The reference cannot be imported as is in this image: quimb, cytoolz, dask, opt_einsum, pydash and
matplotlib are absent and there is no network.  The functions below, however, are pure numpy and only
touch `mps.tensors[k].data`, so they run unmodified once the absent modules are replaced by empty
stubs and the MPS is a duck-typed object:

    computepsi (TNutils.py:627), computepsiprime (:681), the psi'/psi accumulation of learning_step
    (:1396-1413, re-enacted here with the reference's own computepsiprime), generate_sample (:1874,
    its only quimb call `mps.right_canonize()` is a no-op on an already right-canonical input),
    bars_n_stripes (:2243), partial_removal_img (:211), mps_lr (:76).

Everything that needs quimb arithmetic (Tensor.split, canonize, tneinsum3 -> generate_samples, the
cached sweep) CANNOT be executed and stays unpinned (see oracle/born_oracle.py header).

Run (in the build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_reference_golden.py
Writes tests/golden/ref_*.npz.  Inputs are produced with numpy seeds and stored in the fixtures.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


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
    for name in ['cytoolz', 'dask', 'dask.array', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors',
                 'matplotlib.pylab', 'quimb', 'quimb.tensor', 'opt_einsum', 'pydash', 'torchvision',
                 'torchvision.transforms', 'tqdm.notebook']:
        if name in sys.modules:
            continue
        try:
            if name in ('torchvision', 'torchvision.transforms', 'tqdm.notebook'):
                __import__(name)
                continue
        except Exception:
            pass
        _stub(name)
    sys.modules['dask.array'].Array = type('Array', (), {})
    sys.modules['dask'].array = sys.modules['dask.array']
    sys.modules['matplotlib.colors'].LinearSegmentedColormap = _Any
    sys.modules['matplotlib.colors'].ListedColormap = _Any
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['matplotlib'].pylab = sys.modules['matplotlib.pylab']
    sys.modules['quimb'].tensor = sys.modules['quimb.tensor']
    sys.path.insert(0, REF)
    import TNutils  # noqa
    return TNutils


class DuckTensor:
    def __init__(self, data):
        self.data = data


class DuckMPS:
    """What the pure-numpy reference functions need from a quimb MPS."""
    def __init__(self, arrays):
        self.tensors = [DuckTensor(a) for a in arrays]

    def right_canonize(self):   # input is right-canonical already
        return None


def right_canonical_mps(L, D, seed):
    """Random right-canonical, normalised MPS in the reference layout (first (Dr,d), middle
    (Dl,Dr,d), last (Dl,d)).  Built with plain numpy QR so the generator is self-contained."""
    rng = np.random.default_rng(seed)
    mats = [rng.standard_normal((1 if k == 0 else D, 1 if k == L - 1 else D, 2)) for k in range(L)]
    for k in range(L - 1, 0, -1):
        dl, dr, d = mats[k].shape
        q, r = np.linalg.qr(mats[k].reshape(dl, dr * d).T)
        mats[k] = q.T.reshape(-1, dr, d)
        mats[k - 1] = np.einsum('abp,bc->acp', mats[k - 1], r.T)
    mats[0] = mats[0] / np.linalg.norm(mats[0])
    out = []
    for k, t in enumerate(mats):
        if k == 0:
            out.append(np.ascontiguousarray(t[0]))
        elif k == L - 1:
            out.append(np.ascontiguousarray(t[:, 0, :]))
        else:
            out.append(np.ascontiguousarray(t))
    return out


def main():
    T = import_reference()
    out = {}

    # ---- psi / psi' / gradient accumulation ------------------------------------------------
    for tag, (L, D, n, seed) in {'a': (8, 4, 12, 101), 'b': (16, 6, 30, 202), 'c': (28, 10, 16, 303)}.items():
        arrays = right_canonical_mps(L, D, seed)
        mps = DuckMPS(arrays)
        rng = np.random.default_rng(seed + 1)
        imgs = (rng.random((n, L)) < 0.4).astype(np.int64)
        psi = np.array([T.computepsi(mps, img) for img in imgs])
        fix = {'L': L, 'D': D, 'imgs': imgs, 'psi': psi}
        for k, a in enumerate(arrays):
            fix[f'site{k}'] = a
        for k in sorted({0, 1, L // 2, L - 3, L - 2}):
            pp = np.array([T.computepsiprime(mps, img, k) for img in imgs], dtype=object)
            pp = np.stack([np.asarray(x, dtype=np.float64) for x in pp])
            fix[f'psiprime_{k}'] = pp
            # the data-dependent term of learning_step (:1396-1413) needs A; A's quimb contraction
            # is replaced by the equivalent einsum in the CONSUMER of this fixture, not here.
        out[f'ref_psi_{tag}.npz'] = fix

    # ---- single-image sampler ---------------------------------------------------------------
    arrays = right_canonical_mps(20, 8, 404)
    mps = DuckMPS(arrays)
    gens, us = [], []
    for s in range(64):
        np.random.seed(1000 + s)
        gens.append(T.generate_sample(mps))
        np.random.seed(1000 + s)
        us.append([np.random.rand() for _ in range(20)])
    fix = {'samples': np.array(gens), 'uniforms': np.array(us)}
    for k, a in enumerate(arrays):
        fix[f'site{k}'] = a
    out['ref_generate_sample.npz'] = fix

    # ---- datasets / masks / lr --------------------------------------------------------------
    np.random.seed(7)
    bas = np.array(T.bars_n_stripes(30))
    np.random.seed(7)
    draws = np.random.random(4096)
    img = (np.random.default_rng(5).random(28 * 28) < 0.3).astype(np.int64)
    masks = {}
    for axis in (0, 1):
        for half in (0, 1):
            masks[f'mask_a{axis}_h{half}'] = T.partial_removal_img(img, (28, 28), .5, axis, half)
    lr = T.mps_lr(DuckMPS([np.zeros(1)] * 6), 0.5, 0.6, 0.05)
    g = np.random.default_rng(9)
    lr_trace, J_trace = [], []
    for epoch in range(3):
        for site in range(5):
            dn = DuckTensor(g.standard_normal((3, 2, 3, 2)))
            # mps_lr.J does arithmetic on the tensor object itself: feed it a raw ndarray-like
            class _Arr(np.ndarray):
                @property
                def data(self): return np.asarray(self)
            arr = dn.data.view(_Arr)
            J = lr.J(site, arr)
            J_trace.append(np.asarray(J))
        lr.new_epoch()
        lr_trace.append(lr.curr_lr)
    out['ref_misc.npz'] = dict(bas=bas, bas_draws=draws, img=img, lr_trace=np.array(lr_trace),
                               J_trace=np.stack(J_trace), J_seed=9, **masks)

    for name, fix in out.items():
        np.savez_compressed(os.path.join(HERE, name), **fix)
        print('wrote', name, {k: np.asarray(v).shape for k, v in list(fix.items())[:6]})


if __name__ == '__main__':
    main()
