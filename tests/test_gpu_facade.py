"""GPU tests of the TNutils-compatible facade: the reference notebooks' call sequences run unchanged."""
import os

import numpy as np
import pytest

from oracle import born_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def T():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from mps_born_machines_b200 import TNutils
    return TNutils


def test_bars_and_stripes_notebook_flow(T):
    """notebooks/bars_n_stripes.ipynb cells 3-4, 11 (old arity of learning_epoch_cached)."""
    np.random.seed(3)
    bars_mps = T.initialize_mps(4 * 4, 2)
    dataset = np.array(T.bars_n_stripes(30))
    bar_data = np.array([T.tens_picture(img.flatten()) for img in dataset])
    bar_cache = T.left_right_cache(bars_mps, bar_data)
    start = [t.copy() for t in bars_mps.arrays()]
    nlls, _ = T.learning_epoch_cached(bars_mps, bar_data, 16, 0.5, bar_cache, lr_update=lambda x: x * 0.95,
                                      batch_size=30, max_bond=16)
    assert len(nlls) == 16 and all(b <= a + 1e-9 for a, b in zip(nlls, nlls[1:]))
    assert np.log(30) < nlls[-1] < np.log(30) + 0.1          # published: 3.41098 vs ln 30 = 3.4012
    assert bars_mps.bond_sizes() == [2, 4, 8] + [16] * 9 + [8, 4, 2]
    # same trajectory as the oracle in the engine's gauge
    imgs = np.array([d.flatten() for d in dataset])
    cost_o, _ = o.learning_epoch_cached(start, imgs, 16, 0.5, batch_size=30, max_bond=16,
                                        lr_update=lambda x: x * 0.95, gauge='fixed')
    np.testing.assert_allclose(nlls, cost_o, rtol=1e-6)
    # computeNLL on the trained model (centre at site 0 after the sweep)
    assert abs(T.computeNLL(bars_mps, imgs, 0) - nlls[-1]) < 1e-9
    assert abs(T.computeNLL(bars_mps, imgs) - nlls[-1]) < 1e-8
    samples = T.generate_samples(bars_mps, 100)
    assert samples.shape == (100, 16)
    valid = set(map(bytes, o.bars_n_stripes_all(4).astype(np.uint8)))
    frac = np.mean([bytes(s.astype(np.uint8)) in valid for s in samples])
    assert frac > 0.9                                        # a trained BAS model mostly emits valid patterns


def test_momentum_flow_matches_oracle(T):
    """bars_n_stripes.ipynb cell 6: mps_lr annealing + the scalar-mean momentum through update_wrap."""
    np.random.seed(4)
    mps = T.initialize_mps(16, 2)
    start = [t.copy() for t in mps.arrays()]
    imgs = o.bars_n_stripes_all(4)
    data = np.array([T.tens_picture(img) for img in imgs])
    cache = T.left_right_cache(mps, data)
    sav_lr = T.mps_lr(mps, 0.5, 0.6, -np.log(0.95))
    update_wrap = lambda site, div: sav_lr.J(site, div)
    nlls, out_lr = T.learning_epoch_cached(mps, data, 4, sav_lr, cache, lr_update=T.lr_update,
                                           update_wrap=update_wrap, batch_size=30, max_bond=16)
    olr = o.MpsLr(16, 0.5, 0.6, -np.log(0.95))
    def o_update(lr):
        lr.new_epoch(); return lr
    cost_o, _ = o.learning_epoch_cached(start, imgs, 4, olr, batch_size=30, max_bond=16, lr_update=o_update,
                                        update_wrap=lambda k, dn: olr.J(k, dn), gauge='fixed')
    np.testing.assert_allclose(nlls, cost_o, rtol=1e-6)
    assert out_lr.t == 4 and abs(out_lr.curr_lr - 0.5 * 0.95 ** 4) < 1e-12


def test_step_by_step_api(T):
    """learning_step_cached + manual write-back + sequential_update (TNutils.py:1603-1620 as a user would)."""
    np.random.seed(5)
    L, N = 10, 24
    mps = T.initialize_mps(L, 4)
    imgs = (np.random.random((N, L)) < 0.4).astype(int)
    data = np.array([T.tens_picture(img) for img in imgs])
    cache = T.left_right_cache(mps, data, max_bond=8)
    mps_o = [t.copy() for t in mps.arrays()]
    phys = o.embed(imgs)
    oc = o.EnvCache(mps_o, phys, 'pow2')
    nll0 = T.computeNLL_cached(mps, data, cache, 0)
    assert abs(nll0 - o.compute_nll(mps_o, imgs, 0)) < 1e-9
    for index in range(0, 4):
        A = T.learning_step_cached(mps, index, data, 0.05, cache, True, renorm='l2', max_bond=8)
        ref = o.two_site_step(mps_o, index, oc.lenv[index], oc.renv[index + 2], phys[:, index], phys[:, index + 1],
                              0.05, True, 'l2', max_bond=8)
        assert A.tensors[0].data.shape == (ref['left'].shape[1:] if index == 0 else ref['left'].shape)
        if index == 0:
            mps.tensors[index].modify(data=np.transpose(A.tensors[0].data, (1, 0)))
        else:
            mps.tensors[index].modify(data=np.transpose(A.tensors[0].data, (0, 2, 1)))
        mps.tensors[index + 1].modify(data=A.tensors[1].data)
        T.sequential_update(mps, data, cache, index, True)
        o.write_back(mps_o, index, ref['left'], ref['right']); oc.advance_right(mps_o, index)
        np.testing.assert_allclose(A.info['Z'], ref['Z'], rtol=1e-9)
    assert abs(T.computeNLL(mps, imgs) - o.compute_nll(mps_o, imgs)) < 1e-8
    assert abs(T.computepsi(mps, imgs[0]) ** 2 - np.exp(2 * o.logpsi(mps_o, phys[:1])[1][0])) < 1e-9 * T.computepsi(mps, imgs[0]) ** 2


def test_compress(T):
    np.random.seed(6)
    mps = T.initialize_mps(12, 8)
    ref = o.compress([t.copy() for t in mps.arrays()], 4)
    small = T.compress_copy(mps, 4)
    assert small.bond_sizes() == o.bond_sizes(ref) and mps.max_bond() == 8
    assert abs((small @ small) - o.norm2_full(ref)) < 1e-9
    assert abs(small @ mps - np.exp(0.5 * 0) * (T.MPS(ref) @ mps)) < 1e-9          # same truncated state
    T.compress2(mps, 4)
    assert mps.max_bond() == 4


def test_generate_sample_and_reconstruct(T):
    np.random.seed(7)
    mps = T.initialize_mps(16, 6)
    arrays = [t.copy() for t in mps.arrays()]
    np.random.seed(11)
    got = T.generate_sample(mps)
    np.random.seed(11)
    u = [np.random.rand() for _ in range(16)]
    assert got == o.generate_sample(o.right_canonize(arrays), u)
    img = np.array(got)
    corr = T.partial_removal_img(img, (4, 4), .5, 0, 0)
    assert (corr[8:] == -1).all()
    np.random.seed(12)
    rec = T.reconstruct(mps, corr)
    np.random.seed(12)
    u = [np.random.rand() for _ in range(8)]
    assert np.array_equal(rec, o.reconstruct(arrays, corr, u))
    # general masks (known suffix, columns removed, scattered): same bits as the reference procedure
    rng = np.random.default_rng(5)
    masks = [T.partial_removal_img(img, (4, 4), .5, 0, 1), T.partial_removal_img(img, (4, 4), .5, 1, 0),
             T.partial_removal_img(img, (4, 4), .5, 1, 1), np.where(rng.random(16) < 0.5, -1, img), np.full(16, -1)]
    for i, corr in enumerate(masks):
        for rep in range(6):
            np.random.seed(100 + 10 * i + rep)
            rec = T.reconstruct(mps, corr)
            np.random.seed(100 + 10 * i + rep)
            u = [np.random.rand() for _ in range(int((corr == -1).sum()))]
            assert np.array_equal(rec, o.reconstruct(arrays, corr, u)), (i, rep)


def test_training_and_probing(T, tmp_path, monkeypatch):
    """tests/grownup_tests.ipynb flow: periods x epochs with probing, checkpoints and sample generation."""
    monkeypatch.chdir(tmp_path)
    np.random.seed(8)
    imgs = o.bars_n_stripes_all(4)
    mps = T.initialize_mps(16, 2)
    data = np.array([T.tens_picture(img) for img in imgs])
    cache = T.left_right_cache(mps, data)
    costs, samples = T.training_and_probing(2, 2, mps, (4, 4), imgs, data, cache, 30, 0.3, lambda x: x * 0.95,
                                            lambda site, div: div, val_imgs=imgs[:5], period_samples=4, max_bond=8)
    assert len(costs) == 1 + 2 * 2 and costs[-1] < costs[0]
    assert len(samples) == 2 and samples[0].shape == (4, 16)
    mps2, tr, _ = None, None, None
    where = f'T30_L16/'
    assert os.path.isfile(where + '1I1.mps.npz') and os.path.isfile(where + 'gen_imgs/1.npy')
    loaded = T._load_mps(where + '1I1.mps.npz')
    assert abs(T.computeNLL(loaded, imgs, 0) - costs[-1]) < 1e-9          # resume = reload tensors (same NLL)


def test_torched_api_flow(T, tmp_path):
    """notebooks/probing_ex.ipynb cells 7-11: the G3 (torch) call sequence, L2 renormalisation, minibatches."""
    np.random.seed(9)
    imgs = (np.random.random((40, 12)) < 0.4).astype(int)
    _imgs = np.array([T.tens_picture(img) for img in imgs])
    mps = T.initialize_mps(imgs.shape[1], bdim=4)
    start = [t.copy() for t in mps.arrays()]
    inds_dict = {'mps': {}, 'imgs': {}, 'cache': {}}
    torch_mps = T.torchized_mps(mps, inds_dict)
    torch_imgs = T.torchized_imgs(_imgs, inds_dict)
    torch_cache = T.torchized_cache(torch_mps, torch_imgs, inds_dict)
    assert inds_dict['mps'][1] == ('i0', 'i1', 'v1') and len(torch_imgs) == 12
    costs, samples = T.training_and_probing_torched(2, 2, mps, inds_dict, torch_mps, (3, 4), imgs, _imgs, torch_imgs, torch_cache,
                                                    40, 0.05, lr_update=lambda x: x * 0.95, period_samples=3,
                                                    path=str(tmp_path) + '/', max_bond=8)
    assert len(costs) == 5 and costs[-1] < costs[0] and len(samples) == 2
    # full batch + l2 rule is gauge invariant: same curve as the plain oracle
    cost_o, _ = o.learning_epoch_cached(start, imgs, 4, 0.05, renorm='l2', max_bond=8, lr_update=lambda x: x * 0.95)
    np.testing.assert_allclose(costs[1:], cost_o, rtol=1e-6)
    assert abs(T.computeNLL_torched(mps, imgs, torch_mps, torch_imgs, inds_dict) - costs[-1]) < 1e-9
    psi, _ = T.torchized_cache(torch_mps, torch_imgs, inds_dict, computing_psi=True)
    assert abs(-2 * np.log(np.abs(psi)).mean() + np.log(np.sum(mps[0].data ** 2)) - costs[-1]) < 1e-9
    # stochastic epoch: minibatches of 10 through the same engine
    np.random.seed(10)
    c2 = T.cached_stochastic_learning_epoch(mps, imgs, _imgs, 1, 0.02, torch_cache, None, None, None, batch_size=10, max_bond=8)
    assert len(c2) == 1 and np.isfinite(c2[0])


def test_arbitrary_update_wrap_callback(T):
    """update_wrap(site, dNLL) with a user function (here: gradient clipping) goes through the host round trip."""
    np.random.seed(12)
    L, N = 8, 30
    mps = T.initialize_mps(L, 3)
    start = [t.copy() for t in mps.arrays()]
    imgs = (np.random.random((N, L)) < 0.5).astype(int)
    data = np.array([T.tens_picture(img) for img in imgs])
    cache = T.left_right_cache(mps, data, max_bond=6)
    seen = []
    def clip(site, dn):
        seen.append((site, dn.inds))
        return dn * 0.5 if isinstance(dn, T.Tensor) else dn * 0.5
    cost, _ = T.learning_epoch_cached(mps, imgs, data, 2, 0.1, cache, batch_size=N, update_wrap=clip, renorm='l2', max_bond=6)
    cost_o, _ = o.learning_epoch_cached(start, imgs, 2, 0.1, renorm='l2', max_bond=6, update_wrap=lambda k, dn: 0.5 * dn)
    np.testing.assert_allclose(cost, cost_o, rtol=1e-6)
    assert seen[0] == (0, ('v0', 'i1', 'v1')) and seen[1][1] == ('i0', 'v1', 'i2', 'v2') and len(seen) == 2 * 2 * (L - 1)


def test_no_max_bond_lets_the_bond_grow(T):
    """`learning_epoch_cached(...)` / `learning_step_cached(...)` WITHOUT max_bond: the reference's split is limited by
    cutoff=1e-10 only (TNutils.py:1552-1564), so bonds grow past the initial maximum (the engine must not
    clamp at its initial capacity).  Same trajectory and bond profile as the oracle."""
    np.random.seed(9)
    mps = T.initialize_mps(12, 2)
    start = [t.copy() for t in mps.arrays()]
    imgs = (np.random.default_rng(5).random((40, 12)) < 0.4).astype(np.int64)
    data = np.array([T.tens_picture(img) for img in imgs])
    cache = T.left_right_cache(mps, data)
    assert cache.engine.Dcap == 2
    nlls, _ = T.learning_epoch_cached(mps, data, 2, 0.05, cache, batch_size=40)
    cost_o, _ = o.learning_epoch_cached(start, imgs, 2, 0.05, batch_size=40, max_bond=None, gauge='fixed')
    assert max(mps.bond_sizes()) > 2 and cache.engine.Dcap >= max(mps.bond_sizes())
    np.testing.assert_allclose(nlls, cost_o, rtol=1e-6)
    # step-level call without max_bond on a fresh model: chi = number of singular values above the cutoff
    mps2 = T.MPS(start)
    cache2 = T.left_right_cache(mps2, data)
    tn = T.learning_step_cached(mps2, 5, data, 0.05, cache2, True)
    assert tn.tensors[0].data.shape[-1] == tn.info['chi'] > 2


def test_reconstruct_batch_equals_the_per_image_loop(T):
    """facade extension reconstruct_batch == [reconstruct(mps, img) for img in imgs] under the same np.random seed
    (same draws in the same order, TNutils.py:2053-2116), for mixed masks (prefix, general) in one call."""
    np.random.seed(21)
    mps = T.initialize_mps(36, 5)
    rng = np.random.default_rng(22)
    imgs = (rng.random((9, 36)) < 0.4).astype(int)
    corr = []
    for i, img in enumerate(imgs):
        corr.append(T.partial_removal_img(img, (6, 6), .5, i % 2, (i // 2) % 2))
    corr = np.array(corr)
    np.random.seed(5)
    want = np.array([T.reconstruct(mps, c) for c in corr])
    np.random.seed(5)
    got = T.reconstruct_batch(mps, corr)
    assert np.array_equal(got, want)
    assert np.array_equal(got[corr != -1], corr[corr != -1]) and set(np.unique(got)) <= {0, 1}


def test_step_by_step_uploads_only_foreign_tensors(T):
    """learning_step_cached + write-back + sequential_update: tensors the facade handed out itself are not uploaded
    again, anything else (a tensor the user replaced) is; results equal a run that uploads everything."""
    np.random.seed(31)
    imgs = (np.random.default_rng(32).random((50, 10)) < 0.4).astype(np.int64)
    data = np.array([T.tens_picture(img) for img in imgs])
    start = T.initialize_mps(10, 4)                     # (draws from an unseeded generator: one model for all three runs)
    def run(perturb, always_upload):
        mps = start.copy()
        cache = T.left_right_cache(mps, data, max_bond=8)
        eng = cache.engine
        for k in range(0, 6):
            if always_upload:
                eng._handed_out = {}
            tn = T.learning_step_cached(mps, k, data, 0.05, cache, True, max_bond=8, renorm='l2')
            T._write_back(mps, k, tn)
            if perturb and k == 2:                       # the user replaces a tensor between the step and the cache update
                mps.tensors[k + 1].modify(data=mps.tensors[k + 1].data * 1.0)
            T.sequential_update(mps, data, cache, k, True)
        return [t.data.copy() for t in mps.tensors], getattr(eng, '_h2d_bytes', 0)
    a, bytes_a = run(False, False)
    b, bytes_b = run(False, True)
    c, bytes_c = run(True, False)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and all(np.array_equal(x, y) for x, y in zip(a, c))
    assert bytes_a < bytes_c < bytes_b


def test_sequential_update_sees_a_tensor_the_user_changed(T):
    """learning_step_cached enqueues the environment advance of its own result right away; a user who changes the
    tensor before calling sequential_update (TNutils.py:504-526) must get the cache of the CHANGED tensor."""
    np.random.seed(41)
    imgs = (np.random.default_rng(42).random((40, 10)) < 0.4).astype(np.int64)
    data = np.array([T.tens_picture(img) for img in imgs])
    mps = T.initialize_mps(10, 4)
    cache = T.left_right_cache(mps, data, max_bond=8)
    for k in range(0, 4):
        tn = T.learning_step_cached(mps, k, data, 0.05, cache, True, max_bond=8, renorm='l2')
        T._write_back(mps, k, tn)
        if k == 2:
            mps.tensors[k].modify(data=mps.tensors[k].data * 0.5)
        T.sequential_update(mps, data, cache, k, True)
        nll_cached = T.computeNLL_cached(mps, data, cache, k + 1)
        fresh = T.left_right_cache(mps.copy(), data, max_bond=8)
        nll_fresh = T.computeNLL_cached(mps, data, fresh, k + 1)
        assert abs(nll_cached - nll_fresh) < 1e-10, (k, nll_cached, nll_fresh)


def test_staged_download_equals_the_converted_one(T):
    """bm_step(advance & 2) leaves the two new tensors in the reference layout so that get_site is a plain copy
    overlapping the environment advance: same bytes as the conversion path."""
    from mps_born_machines_b200 import BornEngine, mpslib
    L, D, N = 12, 8, 64
    imgs = mpslib.synthetic_images(N, L, seed=5, p_on=0.4)
    mps = mpslib.random_mps(L, D, rng=np.random.default_rng(6), centre=3)
    outs = []
    for staged in (True, False):
        eng = BornEngine(L, D)
        eng.set_data(imgs); eng.set_mps(mps); eng.env_build(3)
        res = []
        for k in range(3, 7):
            eng.step(k, 0.05, True, 'l2', max_bond=D, stage_sites=staged)
            res.append((eng.get_site(k), eng.get_site(k + 1)))
        res.append(tuple(eng.get_mps()))
        outs.append(res)
        eng.close()
    for ra, rb in zip(*outs):
        assert all(np.array_equal(x, y) for x, y in zip(ra, rb))
