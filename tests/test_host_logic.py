"""CPU tests (no GPU): the C-ABI library loads and exports every symbol of include/bm.h, host-side facade
objects, and the multi-rank (world_size 2, gloo) sharding + all-reduce logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import born_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mps_born_machines_b200 import _capi
    hdr = open(os.path.join(ROOT, 'include', 'bm.h')).read()
    declared = set(re.findall(r'\b(bm_[a-z0-9_]+)\s*\(', hdr)) - {'bm_ctx'}
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _capi.lib().bm_version() >= 100


def test_engine_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from mps_born_machines_b200 import BornEngine
    with pytest.raises(RuntimeError):
        BornEngine(8, 4)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: no product module may import, link or execute it (comments that
    cite it are fine)."""
    pkg = os.path.join(ROOT, 'mps_born_machines_b200')
    pat = re.compile(r'^\s*(from\s+oracle|import\s+oracle|from\s+\.+\s*oracle|#include\s+.*oracle)', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f
                assert 'importlib' not in src or 'oracle' not in src, f


def test_facade_objects():
    from mps_born_machines_b200 import TNutils as T
    np.random.seed(0)
    mps = T.initialize_mps(10, 4)
    assert [t.inds for t in mps.tensors][:2] == [('i0', 'v0'), ('i0', 'i1', 'v1')]
    assert mps.tensors[-1].inds == ('i8', 'v9')
    assert mps.bond_sizes() == [4, 4, 4, 4, 4, 4, 4, 4, 2] and mps.max_bond() == 4
    assert abs(mps.norm() - 1.0) < 1e-12 and abs((mps @ mps) - 1.0) < 1e-12
    # canonicalisation preserves the norm and yields isometries (tests/canonicalization.ipynb)
    m2 = mps.copy().canonize(5)
    assert abs((m2 @ m2) - 1.0) < 1e-12 and abs((m2 @ mps) - 1.0) < 1e-12
    A = mps[3] @ mps[4]
    assert A.inds == ('i2', 'v3', 'i4', 'v4')
    assert np.allclose(A.data, o.merge(mps.arrays(), 3))
    assert abs((mps[0] @ mps[0]) - np.sum(mps[0].data ** 2)) < 1e-15
    half = (mps / 2.0)
    assert abs(half.norm() - 0.5) < 1e-12
    # embedding: pixel 1 -> [1,0], pixel 0 -> [0,1], -1 -> None
    tp = T.tens_picture([1, 0, -1])
    assert list(tp[0].data) == [1, 0] and list(tp[1].data) == [0, 1] and tp[2] is None
    assert np.array_equal(T._to_pixels(np.array([T.tens_picture([1, 0, 1]), T.tens_picture([0, 0, 1])])), [[1, 0, 1], [0, 0, 1]])
    assert list(T.one(3).data) == [1, 0] and list(T.zero(3).data) == [0, 1]


def test_facade_datasets_match_reference_golden(golden_dir):
    from mps_born_machines_b200 import TNutils as T
    m = np.load(os.path.join(golden_dir, 'ref_misc.npz'))
    np.random.seed(7)
    assert np.array_equal(np.array(T.bars_n_stripes(30)), m['bas'])
    for axis in (0, 1):
        for half in (0, 1):
            assert np.array_equal(T.partial_removal_img(m['img'], (28, 28), .5, axis, half), m[f'mask_a{axis}_h{half}'])
    with pytest.raises(TypeError):
        T.partial_removal_img(list(m['img']), (28, 28), .5, 0, 0)
    with pytest.raises(ValueError):
        T.partial_removal_img(m['img'], (28, 28), 1.5, 0, 0)
    lr = T.mps_lr(T.MPS([np.zeros((1, 2))] + [np.zeros((1, 1, 2))] * 4 + [np.zeros((1, 2))]), 0.5, 0.6, 0.05)
    g = np.random.default_rng(int(m['J_seed']))
    i = 0
    for epoch in range(3):
        for site in range(5):
            J = lr.J(site, T.Tensor(g.standard_normal((3, 2, 3, 2)), 'abcd'))
            np.testing.assert_allclose(J.data if hasattr(J, 'data') and not isinstance(J, np.ndarray) else J, m['J_trace'][i], rtol=1e-15)
            i += 1
        lr = T.lr_update(lr)
        assert lr.curr_lr == m['lr_trace'][epoch]


def test_shard_bounds():
    from mps_born_machines_b200.parallel import shard_bounds
    for n, w in ((60000, 8), (1000, 3), (5, 8), (30, 2)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert sum(hi - lo for lo, hi in spans) == n


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mps_born_machines_b200.parallel import shard_bounds, allreduce_sum_, replicated_checksum
    rng = np.random.default_rng(5)
    L, D, N, k = 8, 5, 41, 3
    mps = o.canonize(o.random_mps(L, D, rng=rng), k)
    imgs = (rng.random((N, L)) < 0.4).astype(np.int64)
    phys = o.embed(imgs)
    lo, hi = shard_bounds(N, world, rank)
    cache = o.EnvCache(mps, phys[lo:hi], 'pow2')
    A = o.merge(mps, k)
    psi, S = o.psi_and_grad(A, cache.lenv[k], cache.renv[k + 2], phys[lo:hi, k], phys[lo:hi, k + 1])
    lnpsi = np.log(np.abs(psi)) + cache.ls_l[k] + cache.ls_r[k + 2]
    buf = torch.from_numpy(np.concatenate([S.reshape(-1), [lnpsi.sum(), float((psi == 0).sum())]]))
    allreduce_sum_(buf)
    # replicated update on every rank from the identical reduced buffer
    S_tot = buf[:-2].numpy().reshape(A.shape)
    A2 = A - 0.1 * (A / np.sum(A * A) - S_tot / N)
    s = np.linalg.svd((A2 / np.sqrt(np.sum(A2 * A2))).reshape(A.shape[0] * 2, -1), compute_uv=False)
    drift = replicated_checksum(torch.from_numpy(s))
    q.put((rank, buf.numpy().copy(), s, drift))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    rng = np.random.default_rng(5)
    L, D, N, k = 8, 5, 41, 3
    mps = o.canonize(o.random_mps(L, D, rng=rng), k)
    imgs = (rng.random((N, L)) < 0.4).astype(np.int64)
    phys = o.embed(imgs)
    cache = o.EnvCache(mps, phys, 'pow2')
    A = o.merge(mps, k)
    psi, S = o.psi_and_grad(A, cache.lenv[k], cache.renv[k + 2], phys[:, k], phys[:, k + 1])
    lnsum = (np.log(np.abs(psi)) + cache.ls_l[k] + cache.ls_r[k + 2]).sum()
    for rank, buf, s, drift in res:
        np.testing.assert_allclose(buf[:-2].reshape(A.shape), S, rtol=1e-11, atol=1e-12 * np.abs(S).max())
        assert abs(buf[-2] - lnsum) < 1e-9 * abs(lnsum) and buf[-1] == 0
        assert drift == 0.0
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])   # bitwise identical replicas


def test_facade_save_load_and_meanpool(tmp_path, monkeypatch):
    """tests/test_functions.ipynb:598,750: the state is invariant under save/load; meanpool2d vs the reference loop."""
    from mps_born_machines_b200 import TNutils as T
    monkeypatch.chdir(tmp_path)
    np.random.seed(1)
    mps = T.initialize_mps(8, 3)
    train = (np.random.random((5, 8)) < 0.5).astype(int)
    T.save_mps_sets(mps, train, 'run1', test_set=train[:2])
    m2, tr, te = T.load_mps_sets('run1')
    assert all(np.array_equal(a.data, b.data) for a, b in zip(mps.tensors, m2.tensors))
    assert np.array_equal(tr, train) and np.array_equal(te, train[:2])
    where = T.mps_checkpoint(mps, train, train[:2], 0, 3, [1.0], [2.0], path=str(tmp_path) + '/')
    assert os.path.isfile(where + '0I2.mps.npz') and os.path.isfile(where + 'train_set.npy')
    T.mps_checkpoint(mps, train, train[:2], 1, 3, [1.0, 0.9], [2.0, 1.9], path=str(tmp_path) + '/')
    assert not os.path.isfile(where + '0I2.mps.npz') and os.path.isfile(where + '1I2.mps.npz')
    assert np.array_equal(np.load(where + 'trainloss.npy'), [1.0, 0.9])
    imgs = np.random.random((3, 16))
    want = []
    for img in imgs:                       # TNutils.py:2299-2314
        a = img.reshape(4, 4)
        want.append([np.mean([a[c, r], a[c, r + 1], a[c + 1, r], a[c + 1, r + 1]]) for c in range(0, 4, 2) for r in range(0, 4, 2)])
    want = (np.array(want) > 0.3).astype(float)
    assert np.array_equal(T.meanpool2d(imgs, (4, 4)), want)


def test_enum_values_match_header():
    """the Python-side option tables carry the numeric values include/bm.h defines"""
    import re
    from mps_born_machines_b200 import _capi
    hdr = open(os.path.join(os.path.dirname(__file__), '..', 'include', 'bm.h')).read()
    vals = {k: int(v) for k, v in re.findall(r'\b(BM_[A-Z0-9_]+)\s*=\s*(\d+)', hdr)}
    assert _capi.SVD == {'gesvdj': vals['BM_SVD_GESVDJ'], 'gesvd': vals['BM_SVD_GESVD'], 'gram': vals['BM_SVD_GRAM'],
                         'gram_syevd': vals['BM_SVD_GRAM_SYEVD']}
    assert _capi.RENORM == {'max': vals['BM_RENORM_MAX'], 'l2': vals['BM_RENORM_L2']}
    assert _capi.RULE == {'batched': vals['BM_RULE_BATCHED'], 'single': vals['BM_RULE_SINGLE']}


class _StubEngine:
    """Records the engine calls of the facade's step-by-step path (no GPU): the bookkeeping that decides whether
    `sequential_update` still has work to do after `learning_step_cached` enqueued the advance itself."""
    def __init__(self):
        self.calls = []
        self._handed_out = {}
        self._advanced = None
        self._h2d_bytes = 0

    def set_site(self, k, t):
        self._advanced = None
        self.calls.append(('set_site', k))

    def env_build(self, k):
        self.calls.append(('env_build', k))

    def env_advance(self, k, going_right):
        self.calls.append(('env_advance', k, bool(going_right)))


def test_sequential_update_skips_only_the_advance_its_step_already_enqueued():
    """TNutils.py:1603-1620: learning_step_cached -> write-back -> sequential_update.  The facade's step enqueues the
    environment advance right behind its split; sequential_update may skip it ONLY when both tensors are still the
    arrays that step handed out and the direction is the same."""
    from types import SimpleNamespace
    from mps_born_machines_b200 import TNutils as T
    a, b = np.ones((2, 3, 2)), np.ones((3, 2, 2))
    a.flags.writeable = False; b.flags.writeable = False

    def fresh():
        eng = _StubEngine()
        eng._handed_out = {4: a, 5: b}
        eng._advanced = (4, True)
        mps = SimpleNamespace(tensors={4: SimpleNamespace(data=a.transpose(0, 1, 2)), 5: SimpleNamespace(data=b)})
        return eng, mps, SimpleNamespace(engine=eng)

    eng, mps, cache = fresh()                                   # untouched result, same direction: nothing to do
    T.sequential_update(mps, None, cache, 4, True)
    assert eng.calls == [] and eng._advanced is None
    T.sequential_update(mps, None, cache, 4, True)              # a second call is a real cache refresh again
    assert eng.calls == [('env_build', 4), ('env_advance', 4, True)]

    eng, mps, cache = fresh()                                   # other direction: the enqueued advance does not cover it
    T.sequential_update(mps, None, cache, 4, False)
    assert eng.calls == [('env_build', 4), ('env_advance', 4, False)]

    eng, mps, cache = fresh()                                   # the user replaced a tensor: upload + full refresh
    mps.tensors[4].data = a * 0.5
    T.sequential_update(mps, None, cache, 4, True)
    assert eng.calls == [('set_site', 4), ('env_build', 4), ('env_advance', 4, True)]
    assert 4 not in eng._handed_out and eng._h2d_bytes == a.nbytes

    eng, mps, cache = fresh()                                   # another bond: not the step that was advanced
    eng._handed_out = {7: a, 8: b}
    mps.tensors = {7: SimpleNamespace(data=a), 8: SimpleNamespace(data=b)}
    T.sequential_update(mps, None, cache, 7, True)
    assert eng.calls == [('env_build', 7), ('env_advance', 7, True)]
