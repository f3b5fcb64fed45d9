"""CPU tests: the oracle against golden vectors produced by the real reference
(tests/golden/make_reference_golden.py) and against exact linear-algebra properties."""
import os

import numpy as np
import pytest

from oracle import born_oracle as o


def _load_mps(f):
    L = int(f['L']) if 'L' in f.files else sum(1 for k in f.files if k.startswith('site'))
    return [f[f'site{k}'] for k in range(L)]


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_computepsi_matches_reference(golden_dir, tag):
    f = np.load(os.path.join(golden_dir, f'ref_psi_{tag}.npz'))
    mps = _load_mps(f)
    psi = np.array([o.computepsi(mps, im) for im in f['imgs']])
    np.testing.assert_allclose(psi, f['psi'], rtol=1e-13)
    s, lp = o.logpsi(mps, o.embed(f['imgs']))
    np.testing.assert_allclose(s * np.exp(lp), f['psi'], rtol=1e-12)


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_psiprime_matches_reference(golden_dir, tag):
    f = np.load(os.path.join(golden_dir, f'ref_psi_{tag}.npz'))
    mps = _load_mps(f)
    for key in [k for k in f.files if k.startswith('psiprime_')]:
        k = int(key.split('_')[1])
        pp = np.stack([o.computepsiprime(mps, im, k) for im in f['imgs']])
        ref = f[key]
        assert np.array_equal(pp.reshape(ref.shape), ref)


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_grouped_gradient_equals_reference_psiprime_sum(golden_dir, tag):
    """sum_n psi'_n/psi_n built from the REFERENCE's psi' (learning_step, TNutils.py:1396-1413)
    equals the pixel-grouped GEMM form used by the oracle and the CUDA engine."""
    f = np.load(os.path.join(golden_dir, f'ref_psi_{tag}.npz'))
    mps = _load_mps(f)
    phys = o.embed(f['imgs'])
    cache = o.EnvCache(mps, phys, rescale=None)
    for key in [k for k in f.files if k.startswith('psiprime_')]:
        k = int(key.split('_')[1])
        A = o.merge(mps, k)
        pp = f[key].reshape((len(phys),) + A.shape)
        den = np.einsum('nijkl,ijkl->n', pp, A)
        S_ref = (pp / den[:, None, None, None, None]).sum(0)
        psi, S = o.psi_and_grad(A, cache.lenv[k], cache.renv[k + 2], phys[:, k], phys[:, k + 1])
        np.testing.assert_allclose(psi, den, rtol=1e-12)
        np.testing.assert_allclose(psi, f['psi'], rtol=1e-12)
        np.testing.assert_allclose(S, S_ref, rtol=1e-11, atol=1e-9 * np.abs(S_ref).max())
        psi2, S2 = o.psi_and_grad_literal(A, cache.lenv[k], cache.renv[k + 2], phys[:, k], phys[:, k + 1])
        np.testing.assert_allclose(S2, S_ref, rtol=1e-11, atol=1e-9 * np.abs(S_ref).max())


def test_generate_sample_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'ref_generate_sample.npz'))
    mps = [g[f'site{k}'] for k in range(20)]
    for i in range(64):
        assert o.generate_sample(mps, g['uniforms'][i]) == list(g['samples'][i])


def test_misc_matches_reference(golden_dir):
    m = np.load(os.path.join(golden_dir, 'ref_misc.npz'))

    class Replay:
        def __init__(self, d): self.d, self.i = d, 0
        def random(self, n):
            r = self.d[self.i:self.i + n]; self.i += n; return r
    bas = np.array(o.bars_n_stripes(30, 4, Replay(m['bas_draws'])))
    assert np.array_equal(bas, m['bas'])
    all30 = o.bars_n_stripes_all(4)
    assert all30.shape == (30, 16)
    assert set(map(bytes, all30.astype(np.uint8))) == set(map(bytes, m['bas'].reshape(30, 16).astype(np.uint8)))
    for axis in (0, 1):
        for half in (0, 1):
            got = o.partial_removal_img(m['img'], (28, 28), .5, axis, half)
            assert np.array_equal(got, m[f'mask_a{axis}_h{half}'])
    lr = o.MpsLr(6, 0.5, 0.6, 0.05)
    g = np.random.default_rng(int(m['J_seed']))
    i = 0
    for epoch in range(3):
        for site in range(5):
            J = lr.J(site, g.standard_normal((3, 2, 3, 2)))
            np.testing.assert_allclose(J, m['J_trace'][i], rtol=1e-15)
            i += 1
        lr.new_epoch()
        assert lr.curr_lr == m['lr_trace'][epoch]


def test_embedding_convention():
    # pixel 1 -> [1,0] (phys 0), pixel 0 -> [0,1] (phys 1), -1 -> unknown   (TNutils.py:446-450)
    assert list(o.embed(np.array([1, 0, -1]))) == [0, 1, o.UNKNOWN]


def test_canonical_forms_and_norm():
    mps = o.random_mps(9, 5, rng=np.random.default_rng(3))
    n0 = o.norm2_full(mps)
    for c in (0, 4, 8):
        can = o.canonize(mps, c)
        assert abs(o.norm2_full(can) - n0) < 1e-12 * n0          # tests/canonicalization.ipynb cells 9-18
        assert abs(np.sum(can[c] ** 2) - n0) < 1e-12 * n0          # Z = centre . centre
        for k in range(c + 1, 9):
            m = o.site3(can, k)
            np.testing.assert_allclose(np.einsum('abp,cbp->ac', m, m), np.eye(m.shape[0]), atol=1e-12)
        for k in range(0, c):
            m = o.site3(can, k)
            np.testing.assert_allclose(np.einsum('abp,acp->bc', m, m), np.eye(m.shape[1]), atol=1e-12)
    # quimb MPS_rand_state + canonize(0): bonds shrink from the right only
    assert o.bond_sizes(o.canonize(mps, 0)) == [5, 5, 5, 5, 5, 5, 4, 2]


def test_chain_equals_brute_force_enumeration():
    # functions.ipynb:738,759: einsum chain == full contraction; here sum_v psi(v)^2 == Z
    mps = o.canonize(o.random_mps(10, 4, rng=np.random.default_rng(5)), 0)
    allimgs = np.array([[(m >> i) & 1 for i in range(10)] for m in range(1024)])
    _, lp = o.logpsi(mps, o.embed(allimgs))
    assert abs(np.exp(2 * lp).sum() - o.norm2_full(mps)) < 1e-12


def test_split_rules():
    rng = np.random.default_rng(11)
    A = rng.standard_normal((6, 2, 5, 2))
    for absorb in ('left', 'right'):
        left, right, sk, s_all = o.split(A, absorb, max_bond=7)
        assert sk.size == 7 and s_all.size == 10
        rec = np.einsum('apb,bcq->apcq', left, right)
        assert abs(np.linalg.norm(rec - A) - np.sqrt(np.sum(s_all[7:] ** 2))) < 1e-12
        iso = left if absorb == 'right' else right
        if absorb == 'right':
            x = left.reshape(12, 7)
            np.testing.assert_allclose(x.T @ x, np.eye(7), atol=1e-12)
        else:
            x = right.reshape(7, 10)
            np.testing.assert_allclose(x @ x.T, np.eye(7), atol=1e-12)
    # relative cutoff: rank-3 matrix keeps exactly 3 values, at least one is always kept
    B = (rng.standard_normal((12, 3)) @ rng.standard_normal((3, 10))).reshape(6, 2, 5, 2)
    assert o.split(B, 'right')[2].size == 3
    assert o.split(np.zeros((2, 2, 2, 2)) + 0.0, 'right', cutoff=1e-10)[2].size == 1
    assert o.split(A, 'right', cutoff=0.0)[2].size == 10


def test_env_cache_is_scale_invariant_and_logscale_exact():
    mps = o.canonize(o.random_mps(12, 6, rng=np.random.default_rng(1)), 0)
    imgs = (np.random.default_rng(2).random((20, 12)) < 0.5).astype(int)
    phys = o.embed(imgs)
    raw = o.EnvCache(mps, phys, rescale=None)
    _, lp = o.logpsi(mps, phys)
    for mode in ('maxabs', 'pow2'):
        c = o.EnvCache(mps, phys, rescale=mode)
        for k in (0, 3, 10):
            A = o.merge(mps, k)
            psi0, S0 = o.psi_and_grad(A, raw.lenv[k], raw.renv[k + 2], phys[:, k], phys[:, k + 1])
            psi1, S1 = o.psi_and_grad(A, c.lenv[k], c.renv[k + 2], phys[:, k], phys[:, k + 1])
            np.testing.assert_allclose(S1, S0, rtol=1e-10, atol=1e-12 * np.abs(S0).max())
            np.testing.assert_allclose(np.log(np.abs(psi1)) + c.ls_l[k] + c.ls_r[k + 2], lp, rtol=1e-12)
        if mode == 'maxabs':
            assert np.allclose(np.abs(c.lenv[5]).max(axis=1), 1.0)


def test_bars_and_stripes_training_curve():
    """Statistical pin: notebooks/bars_n_stripes.ipynb:334-664 ends at 3.41098 (ln 30 = 3.4012) with
    lr 0.5*0.95^e, batch 30, max_bond 16, init D=2, 16 epochs, 'max' renormalisation (TNutils.py:1542)."""
    imgs = o.bars_n_stripes_all(4)
    mps = o.canonize(o.random_mps(16, 2, rng=np.random.default_rng(0)), 0)
    cost, lr = o.learning_epoch_cached(mps, imgs, 16, 0.5, batch_size=30, max_bond=16,
                                       lr_update=lambda l: l * 0.95)
    assert abs(cost[-1] - 3.41) < 0.01
    assert 6.0 < cost[0] < 6.8 and 4.0 < cost[1] < 4.4
    assert all(b <= a + 1e-9 for a, b in zip(cost, cost[1:]))
    assert o.bond_sizes(mps) == [2, 4, 8] + [16] * 9 + [8, 4, 2]
    assert abs(lr - 0.5 * 0.95 ** 16) < 1e-15
    # the sweep leaves the centre at site 0: Z from site 0 equals the full norm
    assert abs(np.sum(mps[0] ** 2) - o.norm2_full(mps)) < 1e-10 * o.norm2_full(mps)


def test_batched_sampler_distribution_and_rules():
    mps = o.canonize(o.random_mps(8, 4, rng=np.random.default_rng(21)), 0)
    rng = np.random.default_rng(22)
    U = rng.random((8, 20000))
    s = o.generate_samples(mps, U)
    s2 = o.generate_samples(mps, U, rescale=True)
    assert np.array_equal(s, s2)
    # empirical frequencies follow |psi|^2/Z
    allimgs = np.array([[(m >> i) & 1 for i in range(8)] for m in range(256)])
    _, lp = o.logpsi(mps, o.embed(allimgs))
    p = np.exp(2 * lp) / o.norm2_full(mps)
    idx = (s.astype(int) * (1 << np.arange(8))).sum(1)
    freq = np.bincount(idx, minlength=256) / s.shape[0]
    assert np.abs(freq - p).max() < 0.01
    # boundary rule: u == p accepts pixel 1 in the batched sampler (<=, TNutils.py:1840)
    m0 = o.site3(mps, 0)
    p1 = (m0[0, :, 0] @ m0[0, :, 0]) / np.sum(m0 * m0)
    Ub = np.full((8, 1), 0.5); Ub[0, 0] = p1
    assert o.generate_samples(mps, Ub)[0, 0] == 1


def test_reconstruct_conditionals():
    mps = o.canonize(o.random_mps(6, 4, rng=np.random.default_rng(31)), 0)
    rng = np.random.default_rng(32)
    allimgs = np.array([[(m >> i) & 1 for i in range(6)] for m in range(64)])
    _, lp = o.logpsi(mps, o.embed(allimgs))
    p = np.exp(2 * lp); p /= p.sum()
    for img in (np.array([1, 0, 1, -1, -1, -1]), np.array([-1, -1, 0, 1, 1, 0]), np.array([1, -1, 0, -1, 1, -1])):
        unk = np.nonzero(img == -1)[0]
        counts = {}
        n = 4000
        for _ in range(n):
            r = o.reconstruct(mps, img, rng.random(len(unk)))
            assert np.array_equal(r[img != -1], img[img != -1])
            counts[tuple(r[unk])] = counts.get(tuple(r[unk]), 0) + 1
        compat = [i for i in range(64) if all(allimgs[i][j] == img[j] for j in range(6) if img[j] != -1)]
        z = sum(p[i] for i in compat)
        for i in compat:
            assert abs(counts.get(tuple(allimgs[i][unk]), 0) / n - p[i] / z) < 0.04


def test_cached_path_matches_reference(golden_dir):
    """tests/golden/make_reference_golden_cached.py: the reference's own left_right_cache / learning_step_cached /
    learning_epoch_cached / generate_samples code (float32 caches) executed over a stand-in tensor layer."""
    f = np.load(os.path.join(golden_dir, 'ref_cached_path.npz'))
    L = int(f['L'])
    mps = [f[f'site{k}'] for k in range(L)]
    imgs = f['imgs']
    phys = o.embed(imgs)
    c = o.EnvCache(mps, phys, 'maxabs')
    for j in range(L):           # img_cache[n,0,j] = left env of bond j ; img_cache[n,1,j] = right env of bond j+1
        np.testing.assert_allclose(c.lenv[j], f[f'cacheL_{j}'], atol=2e-5)
        np.testing.assert_allclose(c.renv[j + 1], f[f'cacheR_{j}'], atol=2e-5)
    ref = o.two_site_step(mps, 3, c.lenv[3], c.renv[5], phys[:, 3], phys[:, 4], 0.05, True, 'max', max_bond=5)
    assert ref['left'].shape == f['step3_left'].shape and ref['right'].shape == f['step3_right'].shape
    np.testing.assert_allclose(np.einsum('apb,bcq->apcq', ref['left'], ref['right']), f['step3_prod'], atol=5e-5)
    m2 = [t.copy() for t in mps]
    cost, lr = o.learning_epoch_cached(m2, imgs, 3, 0.1, batch_size=len(imgs), max_bond=6, lr_update=lambda x: x * 0.9)
    np.testing.assert_allclose(cost, f['epoch_cost'], rtol=5e-5)
    assert lr == float(f['epoch_lr']) and o.bond_sizes(m2) == list(f['epoch_bonds'])
    trained = [f[f'trained{k}'] for k in range(L)]
    assert np.array_equal(o.generate_samples(trained, f['uniforms']), f['samples'])


def test_eigensolver_algorithm_prototype():
    """numpy restatement of the split's on-chip eigensolver (tools/eigfast_proto.py: Householder tridiagonalisation,
    multisection with Sturm counts, twisted-factorisation vectors, back-transformation, Newton-Schulz polish) against
    LAPACK — the algorithm the CUDA kernels of csrc/bm_eig.cuh implement, checked without a GPU."""
    import importlib.util
    path = os.path.join(os.path.dirname(__file__), '..', 'tools', 'eigfast_proto.py')
    spec = importlib.util.spec_from_file_location('eigfast_proto', path)
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    rng = np.random.default_rng(5)
    n, m = 40, 20
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = 10.0 ** (-4 * np.arange(n) / n)
    G = (Q * s ** 2) @ Q.T
    lam, U = proto.eig_fast(G, m)
    lam_ref = np.linalg.eigvalsh(G)
    np.testing.assert_allclose(lam, lam_ref, rtol=0, atol=1e-14 * lam_ref[-1])
    assert np.abs(U.T @ U - np.eye(m)).max() < 1e-9            # before the polish: eps * |T| / gap
    U = proto.polish(U, 1)
    assert np.abs(U.T @ U - np.eye(m)).max() < 1e-14
    assert np.abs(G @ U - U * lam[::-1][:m]).max() < 1e-14 * lam_ref[-1]
