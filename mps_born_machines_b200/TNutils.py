"""Drop-in facade: the reference's module-level function surface (`from TNutils import *`) on the B200 engine.

Same names, argument meaning and error behaviour as `/root/reference/TNutils.py` for the hot path:

    initialize_mps :386      stater :446 / tens_picture :452        left_right_cache :457
    sequential_update :504   computepsi :627                        computeNLL :872 / computeNLL_cached :892
    learning_step_cached :1510   learning_epoch_cached :1570        compress / compress_copy / compress2 :930-993
    generate_samples :1829   generate_sample :1874                  reconstruct :2053 (any mask)
    bars_n_stripes :2243     partial_removal_img :211               mps_lr :76      zero / one :1822-1827

quimb is not available, so `Tensor` / `MPS` below are minimal duck types exposing what the notebooks touch
(`.tensors`, `[k]`, `.data`, `.inds`, `.shape`, `.modify`, `.bond_sizes()`, `.max_bond()`, `.norm()`, `.H`, `@`,
`/`, `.canonize`, `.left_canonize`, `.right_canonize`, `.normalize`, `.copy`, `.show`).  Every contraction of the
hot path runs in libbm.so on the GPU; host numpy is used only for one-off set-up (random init + QR
canonicalisation) and for the tiny object plumbing the reference API imposes.
"""
import copy as _copy
import math
import os

import numpy as np

from . import mpslib
from .engine import BornEngine, sweep_schedule

__all__ = ['Tensor', 'MPS', 'initialize_mps', 'stater', 'tens_picture', 'left_right_cache', 'sequential_update',
           'computepsi', 'computeNLL', 'computeNLL_cached', 'learning_step_cached', 'learning_epoch_cached',
           'compress', 'compress_copy', 'compress2', 'generate_samples', 'generate_sample', 'reconstruct',
           'bars_n_stripes', 'partial_removal_img', 'mps_lr', 'lr_update', 'zero', 'one', 'ImgCache',
           'torchized_mps', 'torchized_imgs', 'torchized_cache', 'sequential_update_torched', 'learning_epoch_torched',
           'computeNLL_torched', 'training_and_probing_torched', 'cached_stochastic_learning_epoch',
           'training_and_probing', 'mps_checkpoint', 'generate_and_save', 'save_mps_sets', 'load_mps_sets', 'meanpool2d']


# ---------------------------------------------------------------------------------------------------
# minimal quimb-like objects
# ---------------------------------------------------------------------------------------------------
class Tensor:
    def __init__(self, data=1.0, inds=(), tags=None):
        self.data = np.asarray(data, dtype=np.float64)
        self.inds = tuple(inds)
        self.tags = set([tags]) if isinstance(tags, str) else set(tags or ())
        if self.data.ndim != len(self.inds):
            raise ValueError('number of indices does not match the data')

    @property
    def shape(self):
        return self.data.shape

    def modify(self, data=None, inds=None, tags=None):
        if data is not None:
            self.data = np.asarray(data, dtype=np.float64)
        if inds is not None:
            self.inds = tuple(inds)
        if tags is not None:
            self.tags = set([tags]) if isinstance(tags, str) else set(tags)

    def copy(self):
        return Tensor(self.data.copy(), self.inds, set(self.tags))

    @property
    def H(self):
        return self.copy()          # real tensors

    def __matmul__(self, other):
        """contract over shared index names (quimb `@`); returns a float for a full contraction."""
        if isinstance(other, MPS):
            return NotImplemented
        letters = {}
        def sym(i):
            return letters.setdefault(i, chr(ord('a') + len(letters)))
        a = ''.join(sym(i) for i in self.inds)
        b = ''.join(sym(i) for i in other.inds)
        out_inds = [i for i in self.inds if i not in other.inds] + [i for i in other.inds if i not in self.inds]
        out = ''.join(letters[i] for i in out_inds)
        res = np.einsum(f'{a},{b}->{out}', self.data, other.data)
        return float(res) if not out_inds else Tensor(res, out_inds)

    def __truediv__(self, x):
        return Tensor(self.data / x, self.inds, self.tags)

    def __mul__(self, x):
        return Tensor(self.data * x, self.inds, self.tags)

    __rmul__ = __mul__

    def __sub__(self, o):
        return Tensor(self.data - (o.data if isinstance(o, Tensor) else o), self.inds, self.tags)

    def __add__(self, o):
        return Tensor(self.data + (o.data if isinstance(o, Tensor) else o), self.inds, self.tags)

    __radd__ = __add__

    def __repr__(self):
        return f'Tensor(shape={self.shape}, inds={self.inds}, tags={self.tags})'


class TensorNetwork:
    """result of a split: `.tensors[0]` left factor, `.tensors[1]` right factor (quimb returns a TN)."""
    def __init__(self, tensors):
        self.tensors = tuple(tensors)

    def __getitem__(self, i):
        return self.tensors[i]


class MPS:
    """Matrix product state with the reference's index names: site k carries ('i{k-1}','i{k}','v{k}'),
    the first ('i0','v0'), the last ('i{L-2}','v{L-1}')  (TNutils.py:409-422)."""
    cyclic = False

    def __init__(self, arrays):
        L = len(arrays)
        self.tensors = []
        for k, a in enumerate(arrays):
            if k == 0:
                inds = ('i0', 'v0')
            elif k == L - 1:
                inds = (f'i{k - 1}', f'v{k}')
            else:
                inds = (f'i{k - 1}', f'i{k}', f'v{k}')
            self.tensors.append(Tensor(np.array(a, dtype=np.float64), inds, f'I{k}'))
        self.tensors = tuple(self.tensors)

    # -- container protocol
    def __getitem__(self, k):
        return self.tensors[k]

    def __len__(self):
        return len(self.tensors)

    @property
    def L(self):
        return len(self.tensors)

    nsites = L

    def arrays(self):
        return [t.data for t in self.tensors]

    def copy(self):
        return MPS([t.data.copy() for t in self.tensors])

    __copy__ = copy

    @property
    def H(self):
        return self.copy()

    def bond_sizes(self):
        # right bond of site k straight from the stored shapes: (Dr, d) at site 0, (Dl, Dr, d) elsewhere
        return [t.data.shape[0] if (k == 0 and t.data.ndim == 2) else t.data.shape[1]
                for k, t in enumerate(self.tensors[:-1])]

    def max_bond(self):
        return max(self.bond_sizes())

    def _set(self, arrays):
        for t, a in zip(self.tensors, arrays):
            t.modify(data=a)

    def lognorm2(self):
        e = np.ones((1, 1)); ls = 0.0
        for k in range(self.L):
            m = mpslib.site3(self.arrays(), k)
            e = sum(m[:, :, p].T @ e @ m[:, :, p] for p in range(m.shape[2]))
            mx = np.abs(e).max(); e = e / mx; ls += math.log(mx)
        return math.log(e[0, 0]) + ls

    def norm(self):
        """sqrt(<psi|psi>) like quimb (the notebooks divide psi^2 by it, tests/test_functions.ipynb cell 29)."""
        return math.exp(0.5 * self.lognorm2())

    def __matmul__(self, other):
        if other is self or isinstance(other, MPS):
            e = np.ones((1, 1))
            for k in range(self.L):
                a = mpslib.site3(self.arrays(), k); b = mpslib.site3(other.arrays(), k)
                e = sum(a[:, :, p].T @ e @ b[:, :, p] for p in range(a.shape[2]))
            return float(e[0, 0])
        return NotImplemented

    def __truediv__(self, x):
        out = self.copy()
        out.tensors[0].modify(data=out.tensors[0].data / x)
        return out

    def canonize(self, k):
        self._set(mpslib.canonize(self.arrays(), k))
        return self

    def right_canonize(self):
        return self.canonize(0)

    def left_canonize(self):
        return self.canonize(self.L - 1)

    def normalize(self):
        f = math.exp(-self.lognorm2() / (2 * self.L))
        for t in self.tensors:
            t.modify(data=t.data * f)
        return self

    def show(self):
        print(' '.join(str(b) for b in self.bond_sizes()))
        print('o-' * (self.L - 1) + 'o')

    def __repr__(self):
        return f'MPS(L={self.L}, max_bond={self.max_bond()})'


# ---------------------------------------------------------------------------------------------------
# lr schedule (TNutils.py:76-116, verbatim semantics including the scalar-mean "momentum")
# ---------------------------------------------------------------------------------------------------
class mps_lr:
    def __init__(self, mps, lr0, momentum, s_factor):
        self.sites = len(mps.tensors)
        self.lr0 = lr0
        self.curr_lr = self.lr0
        self.momentum = momentum
        self.t = 0
        self.s_factor = s_factor
        self.past_grad = (self.sites - 1) * [0]

    def clear(self):
        self.t = 0

    def compute_lr(self, t):
        return (self.lr0) * np.exp(-self.s_factor * self.t)

    def new_epoch(self):
        self.t = self.t + 1
        self.curr_lr = self.compute_lr(self.t)

    def J(self, left_index, dNLL):
        if self.t > 0:
            J = self.momentum * self.past_grad[left_index] + dNLL
        else:
            J = dNLL
        self.past_grad[left_index] = np.mean(dNLL.data)
        return J


def lr_update(lr):
    lr.new_epoch()
    return lr


# ---------------------------------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------------------------------
def stater(x, i):
    if x in [0, 1]:
        return Tensor([int(x), int(not x)], inds=(f'v{i}',))
    return None


def tens_picture(picture):
    """Converts an array of bits into a list of tensors compatible with a tensor network."""
    return np.array([stater(n, i) for i, n in enumerate(picture)], dtype=object)


_zero, _one = {}, {}


def zero(ind):
    return _zero.setdefault(ind, Tensor([0, 1], inds=(f'v{ind}',)))


def one(ind):
    return _one.setdefault(ind, Tensor([1, 0], inds=(f'v{ind}',)))


def _to_pixels(_imgs):
    """reference-style object array of `stater` tensors (or a plain 0/1/-1 array) -> int array (N,L)."""
    a = np.asarray(_imgs, dtype=object) if not isinstance(_imgs, np.ndarray) else _imgs
    if a.dtype != object:
        return np.asarray(a, dtype=np.int64)
    out = np.empty(a.shape, dtype=np.int64)
    flat_in, flat_out = a.reshape(-1), out.reshape(-1)
    for i, t in enumerate(flat_in):
        flat_out[i] = -1 if t is None else (1 if t.data[0] == 1 else 0)
    return out


def bars_n_stripes(N_samples, dim=4):
    if N_samples > 30:
        N_samples = 30
    samples = []
    while len(samples) < N_samples:
        sample = np.zeros((dim, dim))
        guide = np.random.random(dim + 1)
        if guide[0] <= 0.5:
            sample[guide[1:] <= 0.5, :] = 1
        else:
            sample[:, guide[1:] <= 0.5] = 1
        if not any((sample == x).all() for x in samples):
            samples.append(sample)
    return samples


def partial_removal_img(mnistimg, shape, fraction=.5, axis=0, half=None):
    if len(shape) != 2 or (shape[0] < 1 or shape[1] < 1):
        raise ValueError('The shape of an image needs two positive integer components')
    if [type(mnistimg), type(fraction), type(axis)] != [np.ndarray, float, int]:
        raise TypeError('Input types not valid')
    if mnistimg.shape[0] != shape[0] * shape[1]:
        raise TypeError(f'Input image shape does not match, need (f{shape[0] * shape[1]},)')
    if axis not in [0, 1]:
        raise ValueError('Invalid axis [0,1]')
    if fraction < 0 or fraction > 1:
        raise ValueError('Invalid value for fraction variable (in interval [0,1])')
    corr = np.copy(mnistimg).reshape(shape)
    if half is None:
        half = np.random.randint(2)
    if axis == 0:
        if half == 0:
            corr[int(shape[0] * (1 - fraction)):, :] = -1
        else:
            corr[:int(shape[0] * (1 - fraction)), :] = -1
    else:
        if half == 0:
            corr[:, int(shape[1] * (1 - fraction)):] = -1
        else:
            corr[:, :int(shape[1] * (1 - fraction))] = -1
    return corr.reshape(shape[0] * shape[1],)


# ---------------------------------------------------------------------------------------------------
# model + caches
# ---------------------------------------------------------------------------------------------------
def initialize_mps(Ldim=28 * 28, bdim=30, canonicalize=0):
    """random MPS with every bond = bdim, canonised to `canonicalize` (out of range: skip), indices renamed
    i{k}/v{k}  (TNutils.py:386-424)."""
    rng = np.random.default_rng()
    c = canonicalize if canonicalize in range(Ldim) else None
    arrays = mpslib.random_mps(Ldim, bdim, rng=rng, centre=c)
    return MPS(arrays)


class ImgCache:
    """What `left_right_cache` returns: the device-resident environment stack of all images (one engine).
    Indexing with a minibatch mask (`img_cache[mask]`, TNutils.py:1607) yields a view that remembers the mask."""

    def __init__(self, engine, pixels, mask=None):
        self.engine, self.pixels, self.mask = engine, pixels, mask

    def __len__(self):
        return len(self.pixels) if self.mask is None else len(self.mask)

    def __getitem__(self, mask):
        if isinstance(mask, tuple):
            raise TypeError('environment vectors live on the device; use ImgCache.left(j)/right(j) to inspect them')
        return ImgCache(self.engine, self.pixels, np.asarray(mask))

    def left(self, j):
        return self.engine.env_get(0, j)

    def right(self, j):
        return self.engine.env_get(1, j + 1)


def _engine_for(mps, pixels, max_bond=None):
    d = mps.tensors[0].data.shape[-1]
    cap = max(int(max_bond or 0), mps.max_bond(), 2)
    eng = BornEngine(len(mps.tensors), cap, phys_dim=d)
    eng.set_mps(mps.arrays())
    eng.set_data(pixels)
    return eng


def left_right_cache(mps, _imgs, max_bond=None):
    """TNutils.py:457-481.  `max_bond` (extension) sizes the device buffers for later bond growth; the cache is
    re-created transparently by the learning functions when a step needs more room (an explicit `max_bond=` above the
    capacity, or NO `max_bond`: the reference then lets the bond grow to d*min(Dl, Dr), limited only by `cutoff`)."""
    pixels = _to_pixels(_imgs)
    eng = _engine_for(mps, pixels, max_bond)
    eng.env_build(0)
    return ImgCache(eng, pixels)


def _ensure_capacity(mps, img_cache, max_bond, index=None, dims=None):
    """Make sure the device buffers can hold the bond this step may produce.  With `max_bond=None` the reference's
    `A.split(cutoff=1e-10)` (TNutils.py:1552-1564) keeps up to min(d*Dl, d*Dr) values, so the engine must not clamp
    at its initial capacity: it is re-created with (at least) twice the capacity whenever bond `index` could outgrow
    it.  `dims` = (Dl of site index, Dr of site index+1) when the caller knows them (the sites may live on the
    device only); returns (engine, recreated)."""
    eng = img_cache.engine
    need = max(int(max_bond or 0), 2)
    if not max_bond and index is not None:
        if dims is None:
            a, b = mps.tensors[index].data, mps.tensors[index + 1].data
            dims = (1 if index == 0 else a.shape[0], 1 if index + 1 == len(mps.tensors) - 1 else b.shape[1])
        need = eng.d * min(dims)
    if need <= eng.Dcap:
        return eng, False
    arrays = [t.data for t in mps.tensors]
    need = max(need, max(MPS(arrays).bond_sizes())) if max_bond else max(need, 2 * eng.Dcap)
    eng.close()
    new = BornEngine(len(mps.tensors), need, phys_dim=eng.d)
    new.set_mps(arrays)
    new.set_data(img_cache.pixels)
    img_cache.engine = new
    return new, True


def _same_as_engine(eng, site, data):
    """True when `data` is (a view of) the read-only array the facade handed out for this site after the engine's last
    step on it: the device copy is current and the upload can be skipped.  The arrays are flagged read-only when they
    are handed out, so identity implies unchanged content."""
    ref = getattr(eng, '_handed_out', {}).get(site)
    if ref is None:
        return False
    base = data
    while base is not None and base is not ref:
        base = getattr(base, 'base', None)
    return base is ref and data.size == ref.size


def _upload_site(eng, site, data):
    if not _same_as_engine(eng, site, data):
        eng.set_site(site, data)
        getattr(eng, '_handed_out', {}).pop(site, None)
        eng._advanced = None
        eng._h2d_bytes = getattr(eng, '_h2d_bytes', 0) + int(np.asarray(data).nbytes)


def sequential_update(mps, _imgs, img_cache, site, going_right):
    """TNutils.py:504-526: refresh the cache for ALL images after the two tensors at (site, site+1) changed.
    `learning_step_cached` already enqueued this advance behind its split (it runs while the new tensors are being
    downloaded); when the MPS still holds exactly the tensors that step returned there is nothing left to do."""
    eng = img_cache.engine
    _upload_site(eng, site, mps.tensors[site].data)            # skipped when the tensors are the ones the step just returned
    _upload_site(eng, site + 1, mps.tensors[site + 1].data)
    if getattr(eng, '_advanced', None) == (site, bool(going_right)):
        eng._advanced = None
        return
    eng._advanced = None
    eng.env_build(site)
    eng.env_advance(site, going_right)


def computepsi(mps, img):
    """psi(img) (TNutils.py:627-679), evaluated by the engine's left chain."""
    eng = BornEngine(len(mps.tensors), max(mps.max_bond(), 2), phys_dim=mps.tensors[0].data.shape[-1])
    eng.set_mps(mps.arrays())
    sg, ln = eng.logpsi(np.asarray(img)[None, :])
    eng.close()
    return float(sg[0] * math.exp(ln[0]))


def computeNLL(mps, imgs, canonicalized_index=False):
    """TNutils.py:872-890."""
    imgs = np.asarray(imgs)
    if type(canonicalized_index) == int and 0 <= canonicalized_index <= len(mps.tensors):
        t = mps.tensors[canonicalized_index].data
        logZ = math.log(float(np.sum(t * t)))
    else:
        logZ = mps.lognorm2()
    eng = BornEngine(len(mps.tensors), max(mps.max_bond(), 2), phys_dim=mps.tensors[0].data.shape[-1])
    eng.set_mps(mps.arrays())
    _, ln = eng.logpsi(imgs)
    eng.close()
    return -2 * (ln.sum() / imgs.shape[0]) + logZ


def computeNLL_cached(mps, _imgs, img_cache, index):
    """TNutils.py:892-907: NLL from the cached environments at bond `index`."""
    eng = img_cache.engine
    _upload_site(eng, index, mps.tensors[index].data)
    _upload_site(eng, index + 1, mps.tensors[index + 1].data)
    S = eng.step_grad(index)
    tail = S[-2:].cpu().numpy()
    A = eng.get_merged_at(index)
    return -(2 / len(img_cache)) * tail[0] + math.log(float(np.sum(A * A)))


def _lr_value(lr):
    return getattr(lr, 'curr_lr', lr)       # _pd.get(lr,'curr_lr',lr), TNutils.py:1541


def learning_step_cached(mps, index, _imgs, lr, img_cache, going_right=True,
                         update_wrap=None, renorm='max', **kwargs):
    """TNutils.py:1510-1568.  Returns the two new tensors like quimb's split: `.tensors[0].data` is (Dl,d,chi)
    ((d,chi) at index 0), `.tensors[1].data` is (chi,Dr,d).  `update_wrap(site, dNLL)` callbacks other than the
    identity run on the host (download dNLL, call, upload) — the reference's momentum (`mps_lr.J`) is recognised
    and evaluated on the device instead."""
    max_bond, cutoff = kwargs.get('max_bond'), kwargs.get('cutoff', 1e-10)
    eng, recreated = _ensure_capacity(mps, img_cache, max_bond, index)
    if recreated:
        eng.env_build(index)
    _upload_site(eng, index, mps.tensors[index].data)
    _upload_site(eng, index + 1, mps.tensors[index + 1].data)
    mask = img_cache.mask
    mom_bias = 0.0
    owner = getattr(update_wrap, '__self__', None)
    if update_wrap is not None and not isinstance(owner, mps_lr):
        owner = _find_mps_lr(update_wrap)
    if isinstance(owner, mps_lr):
        if owner.t > 0:
            mom_bias = owner.momentum * owner.past_grad[index]
    generic = update_wrap is not None and not isinstance(owner, mps_lr) and not _is_identity(update_wrap)
    if generic:       # arbitrary callback: host round trip (download dNLL, call, upload the update)
        out = eng.step_with_update_fn(index, _lr_value(lr), lambda site, dn: _as_array(update_wrap(site, _dnll_tensor(site, dn, len(mps.tensors)))).reshape(dn.shape),
                                      going_right, renorm, max_bond, cutoff, mask, advance=False)
    else:
        # advance=True: the `sequential_update` that follows every step in the reference's loop (TNutils.py:1620) is
        # enqueued right behind the split and overlaps the download of the two new tensors
        out = eng.step(index, _lr_value(lr), going_right, renorm, max_bond, cutoff, mask, mom_bias, advance=True,
                       stage_sites=True)
        eng._advanced = (index, bool(going_right))
    if isinstance(owner, mps_lr) and not out['skipped']:
        owner.past_grad[index] = out['mean_dnll']
    if out['skipped']:
        print('RIPPERONI')              # the reference's zero-psi message (TNutils.py:1250)
    a, b = eng.get_site(index, squeeze=False), eng.get_site(index + 1, squeeze=False)
    a.flags.writeable = False; b.flags.writeable = False       # handed out read-only: see _same_as_engine
    if not hasattr(eng, '_handed_out'):
        eng._handed_out = {}
    eng._handed_out[index] = a; eng._handed_out[index + 1] = b
    left = a.transpose(0, 2, 1)          # (Dl,d,chi)
    right = b                            # (chi,Dr,d)
    L = len(mps.tensors)
    lt = Tensor(left[0] if index == 0 else left,
                (f'v{index}', 'bond') if index == 0 else (f'i{index - 1}', f'v{index}', 'bond'))
    rt = Tensor(right[:, 0, :] if index + 1 == L - 1 else right,
                ('bond', f'v{index + 1}') if index + 1 == L - 1 else ('bond', f'i{index + 1}', f'v{index + 1}'))
    tn = TensorNetwork([lt, rt])
    tn.info = out
    return tn


def _as_array(x):
    return x.data if isinstance(x, Tensor) else np.asarray(x)


def _dnll_tensor(site, dn, L):
    """dNLL as the reference hands it to update_wrap: a tensor with inds (i{k-1}, v{k}, i{k+1}, v{k+1}), rank 3 at
    the two ends (TNutils.py:1526)."""
    if site == 0:
        return Tensor(dn[0], (f'v{site}', f'i{site + 1}', f'v{site + 1}'))
    if site == L - 2:
        return Tensor(dn[:, :, 0, :], (f'i{site - 1}', f'v{site}', f'v{site + 1}'))
    return Tensor(dn, (f'i{site - 1}', f'v{site}', f'i{site + 1}', f'v{site + 1}'))


def _is_identity(f):
    try:
        probe = object()
        return f(0, probe) is probe
    except Exception:
        return False


def _find_mps_lr(f):
    """`update_wrap = lambda site, div: sav_lr.J(site, div)` (bars_n_stripes.ipynb cell 6): find the mps_lr."""
    for cell in (getattr(f, '__closure__', None) or ()):
        if isinstance(cell.cell_contents, mps_lr):
            return cell.cell_contents
    g = getattr(f, '__globals__', {})
    for name in getattr(getattr(f, '__code__', None), 'co_names', ()):
        if isinstance(g.get(name), mps_lr):
            return g[name]
    return None


def _write_back(mps, index, tn):
    if index == 0:
        mps.tensors[index].modify(data=np.transpose(tn.tensors[0].data, (1, 0)))
    else:
        mps.tensors[index].modify(data=np.transpose(tn.tensors[0].data, (0, 2, 1)))
    mps.tensors[index + 1].modify(data=tn.tensors[1].data)


def learning_epoch_cached(mps, *args, batch_size=25, update_wrap=None, lr_update=lambda lr: lr, renorm='max',
                          **kwargs):
    """TNutils.py:1570-1631.  Accepts both arities found in the notebooks:
         learning_epoch_cached(mps, val_imgs, _imgs, epochs, initial_lr, img_cache, ...)   (TNutils.py)
         learning_epoch_cached(mps, _imgs, epochs, lr, img_cache, ...)                     (bars_n_stripes.ipynb:118)
    Returns (cost, lr)."""
    if len(args) == 5:
        val_imgs, _imgs, epochs, initial_lr, img_cache = args
    elif len(args) == 4:
        _imgs, epochs, initial_lr, img_cache = args
        val_imgs = None
    else:
        raise TypeError('learning_epoch_cached(mps, [val_imgs,] _imgs, epochs, lr, img_cache, ...)')
    pixels = img_cache.pixels
    val = pixels if val_imgs is None else _to_pixels(val_imgs)
    n = len(pixels)
    batch_size = min(n, batch_size)
    guide = np.arange(n)
    max_bond, cutoff = kwargs.get('max_bond'), kwargs.get('cutoff', 1e-10)
    eng, _ = _ensure_capacity(mps, img_cache, max_bond)
    eng.set_mps(mps.arrays())
    eng._handed_out = {}                 # the sweeps below change every site on the device
    eng.env_build(0)
    owner = None
    generic = False
    if update_wrap is not None and not _is_identity(update_wrap):
        owner = getattr(update_wrap, '__self__', None)
        if not isinstance(owner, mps_lr):
            owner = _find_mps_lr(update_wrap)
        generic = owner is None
    L = len(mps.tensors)
    cost = []
    lr = _copy.copy(initial_lr)
    sched = sweep_schedule(L)
    for epoch in range(epochs):
        print(f'epoch {epoch + 1}/{epochs}')
        going_right = True
        # no momentum feedback between the two visits of a bond, no host callback, bounded bonds: the whole epoch is one
        # library call (BornEngine.sweep); the masks are drawn first, in the order the step-by-step loop draws them
        if not generic and max_bond and (owner is None or owner.t == 0):
            masks = None
            if batch_size < n:
                masks = []
                for _ in sched:
                    np.random.shuffle(guide)
                    masks.append(guide[:batch_size].copy())
                masks = np.stack(masks)
            outs = eng.sweep(sched, [1] * (L - 1) + [0] * (L - 1), _lr_value(lr), renorm, max_bond, cutoff, masks)
            if owner is not None:
                for index, out in zip(sched, outs):
                    if not out['skipped']:
                        owner.past_grad[index] = out['mean_dnll']
            sched_loop = ()
        else:
            sched_loop = sched
        for index in sched_loop:
            mask = None
            if batch_size < n:
                np.random.shuffle(guide)
                mask = guide[:batch_size].copy()
            mom = owner.momentum * owner.past_grad[index] if (owner is not None and owner.t > 0) else 0.0
            if not max_bond:             # unbounded growth (cutoff only): make room before the bond can outgrow the buffers
                dims = (eng.site_dims(index)[0], eng.site_dims(index + 1)[1])
                if eng.d * min(dims) > eng.Dcap:
                    mps._set(eng.get_mps())
                    eng, _ = _ensure_capacity(mps, img_cache, None, index, dims)   # environments are rebuilt by the step
            if generic:
                out = eng.step_with_update_fn(index, _lr_value(lr), lambda site, dn: _as_array(update_wrap(site, _dnll_tensor(site, dn, L))).reshape(dn.shape),
                                              going_right, renorm, max_bond, cutoff, mask)
            else:
                out = eng.step(index, _lr_value(lr), going_right, renorm, max_bond, cutoff, mask, mom)
            if owner is not None and not out['skipped']:
                owner.past_grad[index] = out['mean_dnll']
            if index == L - 2:
                going_right = False
        mps._set(eng.get_mps())
        _, ln = eng.logpsi(val)
        t0 = mps.tensors[0].data
        nll = -2 * (ln.sum() / len(val)) + math.log(float(np.sum(t0 * t0)))     # computeNLL(mps, val_imgs, 0)
        lr = lr_update(lr)
        print('NLL: {} | Baseline: {}'.format(nll, np.log(n)))
        cost.append(nll)
    return cost, lr


# ---------------------------------------------------------------------------------------------------
# compression (TNutils.py:930-993)
# ---------------------------------------------------------------------------------------------------
def _compress_on(mps, max_bond, skip_small):
    eng = BornEngine(len(mps.tensors), max(mps.max_bond(), 2), phys_dim=mps.tensors[0].data.shape[-1])
    eng.set_mps(mps.arrays())
    for index in range(len(mps.tensors) - 2, -1, -1):
        if skip_small and eng.site_dims(index)[1] <= max_bond:
            continue
        eng.split_bond(index, going_right=False, max_bond=max_bond)
    mps._set(eng.get_mps())
    eng.close()
    return mps


def compress(mps, max_bond):
    return _compress_on(mps, max_bond, False)


def compress_copy(mps, max_bond):
    return _compress_on(mps.copy(), max_bond, False)


def compress2(mps, max_bond):
    return _compress_on(mps, max_bond, True)


# ---------------------------------------------------------------------------------------------------
# generation / reconstruction
# ---------------------------------------------------------------------------------------------------
def generate_samples(mps, N):
    """TNutils.py:1829-1872: N samples in lock step; one `np.random.random(N)` per site, in site order, exactly
    like the reference (so identical seeds give identical uniforms)."""
    L = len(mps.tensors)
    U = np.stack([np.random.random(N) for _ in range(L)])
    eng = BornEngine(L, max(mps.max_bond(), 2))
    eng.set_mps(mps.arrays())
    out = eng.sample(N, U=U, rule='batched')
    eng.close()
    return out.astype(np.float64)


def generate_sample(mps):
    """TNutils.py:1874-1995: right-canonise, then sample pixel by pixel (pixel 0 tested first, strict <)."""
    mps.right_canonize()
    L = len(mps.tensors)
    U = np.array([[np.random.rand()] for _ in range(L)])
    eng = BornEngine(L, max(mps.max_bond(), 2))
    eng.set_mps(mps.arrays())
    out = eng.sample(1, U=U, rule='single')
    eng.close()
    return [int(x) for x in out[0]]


def reconstruct(mps, corr_img):
    """TNutils.py:2053-2116.  Known-prefix masks (partial_removal_img axis=0, half=0) take the batched sampler path
    (prefix contracted on the device, remaining sites sampled from the right-canonical tail); any other mask takes
    the general path (right-environment matrices, bm_reconstruct_batch with one image)."""
    return reconstruct_batch(mps, np.asarray(corr_img)[None, :])[0]


def reconstruct_batch(mps, corr_imgs):
    """`[reconstruct(mps, img) for img in corr_imgs]` in one go (extension; the reference reconstructs one image per
    call, e.g. notebooks/mnist.ipynb loops over the test set).  Draws np.random.rand() exactly as that loop would --
    image by image, one number per unknown pixel in site order -- so a seeded run returns the same pictures.
    Images are grouped by mask; each group runs as one batch on the device (known-prefix masks through the batched
    sampler, any other mask through bm_reconstruct_batch)."""
    corr = np.asarray(corr_imgs)
    if corr.ndim != 2 or corr.shape[1] != len(mps.tensors):
        raise ValueError(f'expected (B,{len(mps.tensors)}) images with -1 for the unknown pixels')
    out = corr.copy()
    L = corr.shape[1]
    masks = corr == -1
    draws = [np.array([np.random.rand() for _ in range(int(m.sum()))]) for m in masks]    # the reference's draw order
    groups = {}
    for b, m in enumerate(masks):
        if m.any():
            groups.setdefault(m.tobytes(), []).append(b)
    if not groups:
        return out
    rec = mps.copy().normalize()
    canon = None
    for key, ids in groups.items():
        m = masks[ids[0]]
        start = int(np.nonzero(m)[0][0])
        U = np.stack([draws[b] for b in ids])                 # (B, n_unknown)
        if m[start:].all():                                   # known prefix: batched ancestral sampling of the tail
            if canon is None:
                canon = rec.copy().right_canonize()
                eng_c = BornEngine(L, max(canon.max_bond(), 2))
                eng_c.set_mps(canon.arrays())
            got = eng_c.sample(len(ids), U=np.ascontiguousarray(U.T), rule='single',
                               prefix=corr[ids] if start > 0 else None, start_site=start)
            for i, b in enumerate(ids):
                out[b, start:] = got[i, start:]
        else:
            eng = BornEngine(L, max(rec.max_bond(), 2))
            eng.set_mps(rec.arrays())
            got = eng.reconstruct_batch(corr[ids], U)
            eng.close()
            for i, b in enumerate(ids):
                out[b] = got[i]
    if canon is not None:
        eng_c.close()
    return out


# ---------------------------------------------------------------------------------------------------
# experiment harness, checkpoints (TNutils.py:1682-1733, :2265-2291, :2317-2377) — host-side conveniences.
# quimb pickles are replaced by an open .npz format (site tensors in the reference layout).
# ---------------------------------------------------------------------------------------------------
def _save_mps(mps, path):
    np.savez(path, n_sites=len(mps.tensors), **{f'site{k}': t.data for k, t in enumerate(mps.tensors)})


def _load_mps(path):
    f = np.load(path)
    return MPS([f[f'site{k}'] for k in range(int(f['n_sites']))])


def save_mps_sets(mps, train_set, foldname, test_set=[]):
    """TNutils.py:2265-2278 (mps stored as ./foldname/mps.npz instead of a quimb pickle)."""
    if not os.path.exists('./' + foldname):
        os.makedirs('./' + foldname)
    _save_mps(mps, './' + foldname + '/mps.npz')
    np.save('./' + foldname + '/train_set.npy', train_set)
    if len(test_set) > 0:
        np.save('./' + foldname + '/test_set.npy', test_set)


def load_mps_sets(foldname):
    """TNutils.py:2280-2291."""
    mps = _load_mps('./' + foldname + '/mps.npz')
    train_set = np.load('./' + foldname + '/train_set.npy')
    test_set = np.load('./' + foldname + '/test_set.npy') if os.path.isfile('./' + foldname + '/test_set.npy') else None
    return mps, train_set, test_set


def mps_checkpoint(mps, imgs, val_imgs, period, periods, train_cost, val_cost, path='./'):
    """TNutils.py:2317-2342: folder T{N}_L{L}/, file {period}I{periods-1}.mps(.npz), previous period's file removed,
    trainloss.npy / valloss.npy, and at period 0 the data sets.  No optimiser / RNG state is saved (as in the
    reference): resuming = load the tensors and rebuild the caches with left_right_cache."""
    oldfilename = str(period - 1) + 'I' + str(periods - 1)
    filename = str(period) + 'I' + str(periods - 1)
    foldname = 'T' + str(len(imgs)) + '_L' + str(len(mps.tensors))
    if not os.path.exists(path + foldname):
        os.makedirs(path + foldname)
    if os.path.isfile(path + foldname + '/' + oldfilename + '.mps.npz'):
        os.remove(path + foldname + '/' + oldfilename + '.mps.npz')
    _save_mps(mps, path + foldname + '/' + filename + '.mps.npz')
    np.save(path + foldname + '/trainloss', train_cost)
    np.save(path + foldname + '/valloss', val_cost)
    if period == 0:
        np.save(path + foldname + '/train_set.npy', imgs)
        np.save(path + foldname + '/val_set.npy', val_imgs)
    return path + foldname + '/'


def generate_and_save(mps, period_samples, period, periods, period_epochs, imgs, shape, path='./'):
    """TNutils.py:2344-2377 without the matplotlib figure (SVG output is skipped when matplotlib is unavailable);
    the generated images are saved as gen_imgs/{period}.npy like the reference."""
    gen_imgs = generate_samples(mps, period_samples) if period_samples > 1 else generate_sample(mps)
    if not os.path.exists(path + '/gen_imgs'):
        os.mkdir(path + '/gen_imgs')
    np.save(path + '/gen_imgs/' + str(period) + '.npy', gen_imgs)
    return gen_imgs


def training_and_probing(period_epochs, periods, mps, shape, imgs, _imgs, img_cache, batch_size, lr, lr_update,
                         update_wrap, val_imgs=[], period_samples=0, corrupted_set=None, plot=False, **kwargs):
    """TNutils.py:1682-1733: periods x period_epochs of learning_epoch_cached with NLL probing, checkpoints and
    sample generation after every period.  Returns (train_costs, samples)."""
    train_costs = [computeNLL(mps, imgs, 0)]
    val_costs = []
    if len(val_imgs) > 0:
        val_costs.append(computeNLL(mps, val_imgs, 0))
    samples = []
    for period in range(periods):
        costs, lr = learning_epoch_cached(mps, imgs, _imgs, period_epochs, lr, img_cache, lr_update=lr_update,
                                          update_wrap=update_wrap, batch_size=batch_size, **kwargs)
        train_costs.extend(costs)
        if len(val_imgs) > 0:
            val_costs.append(computeNLL(mps, val_imgs, 0))
        where = mps_checkpoint(mps, imgs, val_imgs, period, periods, train_costs, val_costs)
        if period_samples > 0:
            samples.append(generate_and_save(mps, period_samples, period, periods, period_epochs, imgs, shape, path=where))
    return train_costs, samples


def meanpool2d(npmnist, shape, grayscale_threshold=0.3):
    """TNutils.py:2293-2315: 2x2 mean pooling of flattened images followed by re-binarisation."""
    npmnist = np.asarray(npmnist, dtype=np.float64)
    ds = []
    for img in npmnist:
        a = img.reshape(shape)
        pooled = a.reshape(shape[0] // 2, 2, shape[1] // 2, 2).mean(axis=(1, 3))
        ds.append((pooled.reshape(-1) > grayscale_threshold).astype(np.float64))
    return np.array(ds)


# ---------------------------------------------------------------------------------------------------
# G3 "torched" twins (TNutils.py:995-1378, :1735-1803) and the stochastic-cache epoch (:1633-1675), as shims.
# In the reference these keep float32 torch copies of the sites / images / caches next to the quimb MPS and do the
# contractions with torch.einsum; here the device state lives in the engine (FP64), so the handles are opaque.
# The G3 step renormalises with A/sqrt(A.A) (:1257), so these drivers run the engine with renorm='l2'.
# ---------------------------------------------------------------------------------------------------
class _TorchHandle:
    """what torchized_mps / torchized_imgs return: an opaque handle (indexable per site for len()/debugging)."""
    def __init__(self, kind, payload):
        self.kind, self.payload = kind, payload

    def __len__(self):
        return len(self.payload) if self.kind == 'mps' else self.payload.shape[1]


def torchized_mps(mps, inds_dict):
    """TNutils.py:995-1009: registers the site index names in inds_dict['mps'] like the reference."""
    inds_dict.setdefault('mps', {})
    for site, t in enumerate(mps.tensors):
        inds_dict['mps'][site] = t.inds
    return _TorchHandle('mps', mps)


def torchized_imgs(_imgs, inds_dict):
    """TNutils.py:1011-1027."""
    pixels = _to_pixels(_imgs)
    inds_dict.setdefault('imgs', {})
    for site in range(pixels.shape[1]):
        inds_dict['imgs'][site] = (f'v{site}',)
    return _TorchHandle('imgs', pixels)


def torchized_cache(torch_mps, torch_imgs, inds_dict, computing_psi=False):
    """TNutils.py:1113-1149: left/right environments of all images -> device-resident ImgCache.
    With computing_psi=True the reference returns (psi, cache) from the un-rescaled right chain (:1141-1148)."""
    inds_dict.setdefault('cache', {})
    mps, pixels = torch_mps.payload, torch_imgs.payload
    cache = left_right_cache(mps, pixels)
    if computing_psi:
        sg, ln = cache.engine.logpsi(pixels)
        return sg * np.exp(ln), cache
    return cache


def sequential_update_torched(torch_mps, torch_imgs, torch_cache, site, going_right, inds_dict, rescale=True):
    """TNutils.py:1065-1111."""
    return sequential_update(torch_mps.payload, torch_imgs.payload, torch_cache, site, going_right)


def computeNLL_torched(mps, imgs, torch_mps, torch_imgs, inds_dict):
    """TNutils.py:909-928: NLL of the training images assuming a right-canonical MPS (Z from site 0)."""
    return computeNLL(mps, torch_imgs.payload, 0)


def learning_epoch_torched(mps, imgs, torch_mps, torch_imgs, epochs, initial_lr, torch_cache, inds_dict,
                           batch_size=25, update_wrap=None, lr_update=lambda lr: lr, **kwargs):
    """TNutils.py:1288-1378 (L2 renormalisation, zero-psi steps skipped with the reference's message)."""
    return learning_epoch_cached(mps, imgs, torch_imgs.payload, epochs, initial_lr, torch_cache, batch_size=batch_size,
                                 update_wrap=update_wrap, lr_update=lr_update, renorm='l2', **kwargs)


def training_and_probing_torched(period_epochs, periods, mps, inds_dict, torch_mps, shape, imgs, _imgs, torch_imgs,
                                 torch_cache, batch_size, lr, update_wrap=None, lr_update=lambda lr: lr, val_imgs=[],
                                 period_samples=0, corrupted_set=None, plot=False, path='./', **kwargs):
    """TNutils.py:1735-1803."""
    train_costs = [computeNLL(mps, imgs, 0)]
    val_costs = [computeNLL(mps, val_imgs, 0)] if len(val_imgs) > 0 else []
    samples = []
    for period in range(periods):
        costs, lr = learning_epoch_torched(mps, imgs, torch_mps, torch_imgs, period_epochs, lr, torch_cache, inds_dict,
                                           batch_size=batch_size, update_wrap=update_wrap, lr_update=lr_update, **kwargs)
        train_costs.extend(costs)
        if len(val_imgs) > 0:
            val_costs.append(computeNLL(mps, val_imgs, 0))
        where = mps_checkpoint(mps, imgs, val_imgs, period, periods, train_costs, val_costs, path)
        if period_samples > 0:
            samples.append(generate_and_save(mps, period_samples, period, periods, period_epochs, imgs, shape, path=where))
    return train_costs, samples


def cached_stochastic_learning_epoch(mps, val_imgs, _imgs, epochs, lr, img_cache, last_dirs=None, last_sites=None,
                                     last_epochs=None, batch_size=25, **kwargs):
    """TNutils.py:1633-1675.  The reference refreshes per-image caches lazily (last_dirs/last_sites/last_epochs
    bookkeeping, :586-625) to save CPU time; the device keeps every image's environments current, so the bookkeeping
    arrays are accepted and ignored.  Returns the list of epoch NLLs like the reference."""
    cost, _ = learning_epoch_cached(mps, val_imgs, _imgs, epochs, lr, img_cache, batch_size=batch_size, **kwargs)
    return cost
