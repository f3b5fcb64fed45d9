"""CPU oracle for the MPS Born machine hot path (numpy float64).

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement of the reference algorithm
(`/root/reference/TNutils.py`) for the two-site DMRG-style training sweep and the ancestral
samplers.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it; the product path (`mps_born_machines_b200`) never does.

Parity pin status: PINNED except for the truncation rule of the un-vendored quimb `Tensor.split`.
  * Pinned against the real reference code executed here (generators under tests/golden/, fixtures ref_*.npz):
    - pure-numpy functions run unmodified with the absent imports stubbed (`make_reference_golden.py`):
      `computepsi`, `computepsiprime`, the psi'/psi accumulation of `learning_step`, `generate_sample`,
      `bars_n_stripes`, `partial_removal_img`, `mps_lr`  (psi to 4e-16, psi' and sampled bits identical);
    - the cached (G2) path and the batched sampler run unmodified over a stand-in tensor layer
      (`make_reference_golden_cached.py`: named-index einsum, np.linalg.svd): `tneinsum3`, `left_right_cache`,
      `sequential_update`, `arr_psi_primed_cache`, `learning_step_cached`, `learning_epoch_cached`, `computeNLL`,
      `generate_samples`  (float32 caches in the reference -> agreement 2e-6 on caches, 6e-6 on a 3-epoch NLL curve,
      identical bond profile, identical 2400 sampled bits).
  * UNPINNED: the truncation rule inside `Tensor.split` (quimb 1.3.0 defaults: cutoff=1e-10 relative to s[0], keep
    >= 1, cap at max_bond, no renormalisation) and quimb's `canonize` / `MPS_rand_state`, restated from quimb's
    published behaviour (quimb is absent from /root/reference and not installable offline).  Validated through exact
    linear-algebra properties and the published bars-and-stripes NLL curve
    (`notebooks/bars_n_stripes.ipynb:334-664`), a statistical pin.

Conventions (TNutils.py:386-455):
  * site tensors in quimb 'lrp' layout: first (Dr,d); middle (Dl,Dr,d); last (Dl,d)
  * pixel 1 -> [1,0] (physical index 0), pixel 0 -> [0,1] (physical index 1): phys = 1 - pixel
  * merged tensor A index order (i{k-1}, v{k}, i{k+1}, v{k+1}) = (Dl,d,Dr,d)  (TNutils.py:1388)
  * environments: Lenv[j] = sites 0..j-1 contracted with the image (img_cache[:,0,j]),
    Renv[j] = sites j..L-1 contracted (img_cache[:,1,j-1]);  Lenv[0] = Renv[L] = [1].
"""
from __future__ import annotations

import numpy as np

UNKNOWN = 255  # physical index used for pixel == -1 (stater returns None, TNutils.py:446-450)


# ----------------------------------------------------------------------------------------------
# data embedding and datasets
# ----------------------------------------------------------------------------------------------
def embed(imgs, d=2):
    """pixels -> physical indices.  TNutils.py:446-455 (`stater`/`tens_picture`): x -> [x, not x],
    i.e. pixel 1 selects component 0.  Generalised to d levels as phys = d-1-level."""
    imgs = np.asarray(imgs)
    out = np.full(imgs.shape, UNKNOWN, dtype=np.uint8)
    known = imgs >= 0
    out[known] = (d - 1 - imgs[known].astype(np.int64)).astype(np.uint8)
    return out


def bars_n_stripes_all(dim=4):
    """All 2*2^dim - 2 distinct bars-and-stripes images (30 for dim=4): the set
    `bars_n_stripes(30)` (TNutils.py:2243-2263) converges to, in a deterministic order."""
    out = []
    for orient in (0, 1):
        for m in range(2 ** dim):
            bits = np.array([(m >> i) & 1 for i in range(dim)], dtype=np.float64)
            img = np.zeros((dim, dim))
            if orient == 0:
                img[bits == 1, :] = 1
            else:
                img[:, bits == 1] = 1
            if not any((img == x).all() for x in out):
                out.append(img)
    return np.array([x.reshape(-1) for x in out])


def bars_n_stripes(n_samples, dim=4, rng=None):
    """TNutils.py:2243-2263, with the global np.random replaced by an explicit generator-like
    object exposing `.random(n)` (pass `np.random` to reproduce the reference draw for draw)."""
    rng = np.random if rng is None else rng
    n_samples = min(n_samples, 30)
    samples = []
    while len(samples) < n_samples:
        sample = np.zeros((dim, dim))
        guide = rng.random(dim + 1)
        if guide[0] <= 0.5:
            sample[guide[1:] <= 0.5, :] = 1
        else:
            sample[:, guide[1:] <= 0.5] = 1
        if not any((sample == x).all() for x in samples):
            samples.append(sample)
    return samples


def partial_removal_img(img, shape, fraction=.5, axis=0, half=0):
    """TNutils.py:211-254 (argument checks omitted; `half` must be given)."""
    c = np.copy(img).reshape(shape)
    if axis == 0:
        if half == 0:
            c[int(shape[0] * (1 - fraction)):, :] = -1
        else:
            c[:int(shape[0] * (1 - fraction)), :] = -1
    else:
        if half == 0:
            c[:, int(shape[1] * (1 - fraction)):] = -1
        else:
            c[:, :int(shape[1] * (1 - fraction))] = -1
    return c.reshape(-1)


def synthetic_images(n, L=784, seed=1234, n_proto=64, p_on=0.15, p_flip=0.05):
    """Synthetic binarised images: 64 prototypes, 5% bit flips."""
    rng = np.random.default_rng(seed)
    protos = rng.random((n_proto, L)) < p_on
    c = rng.integers(n_proto, size=n)
    flips = rng.random((n, L)) < p_flip
    return (protos[c] ^ flips).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# MPS container helpers (plain list of ndarrays in the reference layout)
# ----------------------------------------------------------------------------------------------
def site3(mps, k):
    """(Dl,Dr,d) view of site k (Dl=1 at the first site, Dr=1 at the last)."""
    t = mps[k]
    if t.ndim == 3:
        return t
    if k == 0:
        return t.reshape(1, t.shape[0], t.shape[1])
    return t.reshape(t.shape[0], 1, t.shape[1])


def from_site3(mats):
    """inverse of site3 for a whole chain."""
    L = len(mats)
    out = []
    for k, t in enumerate(mats):
        if k == 0:
            out.append(np.ascontiguousarray(t.reshape(t.shape[1], t.shape[2])))
        elif k == L - 1:
            out.append(np.ascontiguousarray(t.reshape(t.shape[0], t.shape[2])))
        else:
            out.append(np.ascontiguousarray(t))
    return out


def bond_sizes(mps):
    return [site3(mps, k).shape[1] for k in range(len(mps) - 1)]


def random_mps(L, bond_dim, d=2, rng=None, normalize=True):
    """Like qtn.MPS_rand_state(L, bond_dim) (TNutils.py:394): i.i.d. normal entries, every bond
    = bond_dim, globally normalised.  (quimb's own RNG cannot be reproduced; parity tests inject
    identical tensors instead.)"""
    rng = np.random.default_rng(0) if rng is None else rng
    mats = []
    for k in range(L):
        dl = 1 if k == 0 else bond_dim
        dr = 1 if k == L - 1 else bond_dim
        mats.append(rng.standard_normal((dl, dr, d)) / np.sqrt(dr * d))
    mps = from_site3(mats)
    if normalize:
        mps = normalize_mps(mps)
    return mps


def _lq_right(m3):
    """(Dl,Dr,d) -> (Lfac (Dl,r), Q (r,Dr,d)) with Q a right isometry: sum_{b,p} Q[a,b,p]Q[a',b,p]=I."""
    dl, dr, d = m3.shape
    q, r = np.linalg.qr(m3.reshape(dl, dr * d).T)   # (dr*d, r), (r, dl)
    return r.T, q.T.reshape(-1, dr, d)


def _qr_left(m3):
    """(Dl,Dr,d) -> (Q (Dl,r,d) left isometry, Rfac (r,Dr))."""
    dl, dr, d = m3.shape
    x = m3.transpose(0, 2, 1).reshape(dl * d, dr)
    q, r = np.linalg.qr(x)
    return q.reshape(dl, d, -1).transpose(0, 2, 1), r


def canonize(mps, centre):
    """quimb `MatrixProductState.canonize(centre)` (called at TNutils.py:398): QR sweeps so that
    sites < centre are left isometries and sites > centre right isometries.  Returns a new list."""
    L = len(mps)
    m = [site3(mps, k).copy() for k in range(L)]
    for k in range(0, centre):
        q, r = _qr_left(m[k])
        m[k] = q
        m[k + 1] = np.tensordot(r, m[k + 1], axes=([1], [0]))
    for k in range(L - 1, centre, -1):
        lf, q = _lq_right(m[k])
        m[k] = q
        m[k - 1] = np.tensordot(m[k - 1], lf, axes=([1], [0])).transpose(0, 2, 1)
    return from_site3(m)


def right_canonize(mps):
    """quimb `right_canonize()` (TNutils.py:1929): orthogonality centre at site 0."""
    return canonize(mps, 0)


def left_canonize(mps):
    return canonize(mps, len(mps) - 1)


def lognorm2(mps):
    """ln <mps|mps> by transfer-matrix contraction with per-site rescaling (safe for long chains)."""
    e = np.ones((1, 1))
    ls = 0.0
    for k in range(len(mps)):
        m = site3(mps, k)
        e = sum(m[:, :, p].T @ e @ m[:, :, p] for p in range(m.shape[2]))
        mx = np.abs(e).max()
        e = e / mx
        ls += np.log(mx)
    return float(np.log(e[0, 0]) + ls)


def norm2_full(mps):
    """<mps|mps> (the reference's `mps @ mps`, TNutils.py:884)."""
    return float(np.exp(lognorm2(mps)))


def normalize_mps(mps):
    """quimb `normalize()` (TNutils.py:2000,2057): <mps|mps> = 1.  The factor is spread evenly over the
    sites (only the global scale matters for every quantity we compare)."""
    f = np.exp(-lognorm2(mps) / (2 * len(mps)))
    return [t * f for t in mps]


# ----------------------------------------------------------------------------------------------
# psi chain, NLL
# ----------------------------------------------------------------------------------------------
def computepsi(mps, img):
    """Literal restatement of TNutils.py:627-679 (left-to-right chain, no rescaling)."""
    d = mps[0].shape[-1]
    ph = embed(np.asarray(img), d)
    c = site3(mps, 0)[0, :, ph[0]]
    for k in range(1, len(mps)):
        c = c @ site3(mps, k)[:, :, ph[k]]
    return float(c[0])


def computepsiprime(mps, img, k):
    """psi'(v) with sites k,k+1 open, index order of the merged tensor (TNutils.py:681-853):
    (Dl,d,Dr,d), one-hot in the two physical legs."""
    d = mps[0].shape[-1]
    L = len(mps)
    ph = embed(np.asarray(img), d)
    lv = np.ones(1)
    for j in range(k):
        lv = lv @ site3(mps, j)[:, :, ph[j]]
    rv = np.ones(1)
    for j in range(L - 1, k + 1, -1):
        rv = site3(mps, j)[:, :, ph[j]] @ rv
    out = np.zeros((lv.size, d, rv.size, d))
    out[:, ph[k], :, ph[k + 1]] = np.outer(lv, rv)
    return out


def logpsi(mps, phys):
    """Stable version of the computepsi chain for a batch: returns (sign, ln|psi|) per sample.
    Same arithmetic as TNutils.py:627-679 with a power-of-two rescale per site (exact)."""
    phys = np.asarray(phys)
    n = phys.shape[0]
    d = mps[0].shape[-1]
    v = np.ones((n, 1))
    ls = np.zeros(n)
    for k in range(len(mps)):
        m = site3(mps, k)
        new = np.empty((n, m.shape[1]))
        for p in range(d):
            idx = np.nonzero(phys[:, k] == p)[0]
            if idx.size:
                new[idx] = v[idx] @ m[:, :, p]
        mx = np.abs(new).max(axis=1)
        e = np.where(mx > 0, np.frexp(np.where(mx > 0, mx, 1.0))[1], 0)
        v = np.ldexp(new, -e[:, None])
        ls += e * np.log(2.0)
    psi = v[:, 0]
    with np.errstate(divide='ignore'):
        return np.sign(psi), np.log(np.abs(psi)) + ls


def compute_nll(mps, imgs, canonicalized_index=False):
    """TNutils.py:872-890: -2/N sum ln|psi| + ln Z, Z from the centre tensor if given."""
    d = mps[0].shape[-1]
    if (type(canonicalized_index) is int) and 0 <= canonicalized_index <= len(mps):
        t = mps[canonicalized_index]
        Z = float(np.sum(t * t))
    else:
        Z = norm2_full(mps)
    _, lp = logpsi(mps, embed(imgs, d))
    return float(-2.0 * lp.sum() / len(imgs) + np.log(Z))


# ----------------------------------------------------------------------------------------------
# environments (TNutils.py:457-481, 504-526)
# ----------------------------------------------------------------------------------------------
def _grouped_apply_right(v, m3, col):
    """new[n,:] = v[n,:] @ m3[:,:,col[n]]  (pixel-value grouped GEMMs)."""
    new = np.empty((v.shape[0], m3.shape[1]))
    for p in range(m3.shape[2]):
        idx = np.nonzero(col == p)[0]
        if idx.size:
            new[idx] = v[idx] @ m3[:, :, p]
    return new


def _grouped_apply_left(v, m3, col):
    """new[n,:] = m3[:,:,col[n]] @ v[n,:]."""
    new = np.empty((v.shape[0], m3.shape[0]))
    for p in range(m3.shape[2]):
        idx = np.nonzero(col == p)[0]
        if idx.size:
            new[idx] = v[idx] @ m3[:, :, p].T
    return new


def _rescale(new, mode):
    """'maxabs' = the reference (TNutils.py:464-465, 513-514); 'pow2' = exponent of max-abs
    (what the CUDA engine does; differs only by a positive per-sample factor); None = no rescale.
    Returns (scaled, ln(scale))."""
    if mode is None:
        return new, np.zeros(new.shape[0])
    mx = np.abs(new).max(axis=1)
    if mode == 'maxabs':
        return new / mx[:, None], np.log(mx)
    e = np.where(mx > 0, np.frexp(np.where(mx > 0, mx, 1.0))[1], 0)
    return np.ldexp(new, -e[:, None]), e * np.log(2.0)


class EnvCache:
    """Left/right environments for all samples: `left_right_cache` (TNutils.py:457-481).
    lenv[j]: (N, D_j) for j=0..L-1 ; renv[j]: (N, D_j) for j=1..L (D_0=D_L=1).
    ls_l / ls_r accumulate ln(scale) so that ln|psi_n| = ln|<A,psi'_n>| + ls_l[k][n] + ls_r[k+2][n]."""

    def __init__(self, mps, phys, rescale='maxabs'):
        L = len(mps)
        n = phys.shape[0]
        self.L, self.n, self.mode, self.phys = L, n, rescale, phys
        self.lenv = [None] * (L + 1)
        self.renv = [None] * (L + 1)
        self.ls_l = [None] * (L + 1)
        self.ls_r = [None] * (L + 1)
        self.lenv[0] = np.ones((n, 1)); self.ls_l[0] = np.zeros(n)
        self.renv[L] = np.ones((n, 1)); self.ls_r[L] = np.zeros(n)
        for k in range(L - 1):
            self.advance_right(mps, k)
        for k in range(L - 1, 0, -1):
            self.advance_left(mps, k)

    def advance_right(self, mps, k):
        """lenv[k+1] from lenv[k] and site k (sequential_update going_right, TNutils.py:507-516)."""
        new = _grouped_apply_right(self.lenv[k], site3(mps, k), self.phys[:, k])
        new, ls = _rescale(new, self.mode)
        self.lenv[k + 1] = new
        self.ls_l[k + 1] = self.ls_l[k] + ls

    def advance_left(self, mps, k):
        """renv[k] from renv[k+1] and site k (sequential_update going left, TNutils.py:517-526)."""
        new = _grouped_apply_left(self.renv[k + 1], site3(mps, k), self.phys[:, k])
        new, ls = _rescale(new, self.mode)
        self.renv[k] = new
        self.ls_r[k] = self.ls_r[k + 1] + ls


# ----------------------------------------------------------------------------------------------
# two-site step (TNutils.py:1380-1445 arithmetic, :1510-1568 cached organisation)
# ----------------------------------------------------------------------------------------------
def merge(mps, k):
    """A = M_k . M_{k+1}, (Dl,d,Dr,d)  (TNutils.py:1388, :1526)."""
    return np.tensordot(site3(mps, k), site3(mps, k + 1), axes=([1], [0]))    # (a,p,c,q), BLAS-backed


def psi_and_grad(A, lv, rv, pk, pk1):
    """psi_n = <A, psi'_n>  and  S = sum_n psi'_n / psi_n  (TNutils.py:1534-1536) in the
    pixel-pair grouped GEMM form: psi = rowdot(L_g A[:,p,:,q], R_g); S[:,p,:,q] = L_g^T diag(1/psi) R_g."""
    dl, d, dr, _ = A.shape
    n = lv.shape[0]
    psi = np.empty(n)
    S = np.zeros_like(A)
    for p in range(d):
        for q in range(d):
            idx = np.nonzero((pk == p) & (pk1 == q))[0]
            if not idx.size:
                continue
            lg, rg = lv[idx], rv[idx]
            ps = np.einsum('nb,nb->n', lg @ A[:, p, :, q], rg)
            psi[idx] = ps
            S[:, p, :, q] = lg.T @ (rg / ps[:, None])
    return psi, S


def psi_and_grad_literal(A, lv, rv, pk, pk1):
    """Same as psi_and_grad, written sample by sample with the explicit psi' outer product
    (TNutils.py:1399-1411).  Slow; used to validate the grouped form."""
    S = np.zeros_like(A)
    psi = np.empty(lv.shape[0])
    for n in range(lv.shape[0]):
        pp = np.zeros_like(A)
        pp[:, pk[n], :, pk1[n]] = np.outer(lv[n], rv[n])
        den = np.einsum('ijkl,ijkl', A, pp)
        psi[n] = den
        S = S + pp / den
    return psi, S


def gauge_signs(U):
    """Deterministic sign convention for singular vectors used by the CUDA engine (k_gauge_signs):
    column b is flipped so that sum_i w_i U[i,b] > 0 with the fixed weights w_i = 1 + frac(phi*(i+1)).
    The reference keeps whatever LAPACK returns; its `A/A.data.max()` step (TNutils.py:1542) makes the
    trajectory depend on these signs, so trajectory-level parity needs both sides in one gauge."""
    i = np.arange(1, U.shape[0] + 1, dtype=np.float64)
    w = 1.0 + np.modf(0.6180339887498949 * i)[0]
    sg = np.sign(w @ U)
    sg[sg == 0] = 1.0
    return sg


def split(A, absorb, max_bond=None, cutoff=1e-10, gauge=None):
    """quimb 1.3.0 `Tensor.split(left_inds=[i{k-1},v{k}], absorb=..., max_bond=, cutoff=)` with its
    defaults method='svd', cutoff_mode='rel', renorm=None (call sites TNutils.py:1434-1441,
    :1557-1564).  UNPINNED restatement (quimb is not vendored):
      thin SVD of the (Dl*d) x (Dr*d) matricisation; keep s_i > cutoff*s_0 (at least one);
      then cap at max_bond; no renormalisation of kept values;
      absorb='right' -> (U, diag(s) Vh); absorb='left' -> (U diag(s), Vh).
    Returns (left (Dl,d,chi), right (chi,Dr,d), s_kept, s_all)."""
    dl, d, dr, d2 = A.shape
    U, s, Vh = np.linalg.svd(A.reshape(dl * d, dr * d2), full_matrices=False)
    if cutoff > 0.0:
        chi = int(np.sum(s > cutoff * s[0]))
        chi = max(chi, 1)
        if max_bond is not None and max_bond > 0:
            chi = min(chi, max_bond)
    elif max_bond is not None and max_bond > 0:
        chi = min(max_bond, s.size)
    else:
        chi = s.size
    U, sk, Vh = U[:, :chi], s[:chi], Vh[:chi]
    if gauge == 'fixed':
        sg = gauge_signs(U)
        U, Vh = U * sg[None, :], Vh * sg[:, None]
    if absorb == 'right':
        Vh = sk[:, None] * Vh
    elif absorb == 'left':
        U = U * sk[None, :]
    else:
        raise ValueError(absorb)
    return U.reshape(dl, d, chi), Vh.reshape(chi, dr, d2), sk, s


def write_back(mps, k, left, right):
    """TNutils.py:1612-1617: left factor (Dl,d,chi) stored transposed as (Dl,chi,d) ((chi,d) at
    site 0); right factor (chi,Dr,d) stored as is ((chi,d) at the last site)."""
    L = len(mps)
    lt = left.transpose(0, 2, 1)
    mps[k] = np.ascontiguousarray(lt[0] if k == 0 else lt)
    mps[k + 1] = np.ascontiguousarray(right[:, 0, :] if k + 1 == L - 1 else right)


def two_site_step(mps, k, lv, rv, pk, pk1, lr, going_right=True, renorm='max',
                  max_bond=None, cutoff=1e-10, update_wrap=None, batch=None, gauge=None):
    """One `learning_step_cached` (TNutils.py:1510-1568) / `learning_step` (:1380-1445).
    lv, rv: environments of the (mini)batch; returns dict with the new factors and diagnostics.
    renorm='max' is :1542 (signed max), 'l2' is :1419."""
    A = merge(mps, k)
    Z = float(np.sum(A * A))
    psi, S = psi_and_grad(A, lv, rv, pk, pk1)
    B = lv.shape[0] if batch is None else batch
    dNLL = A / Z - S / B
    upd = dNLL if update_wrap is None else update_wrap(k, dNLL)
    A2 = A - lr * upd
    if renorm == 'max':
        A2 = A2 / A2.max()
    elif renorm == 'l2':
        A2 = A2 / np.sqrt(np.sum(A2 * A2))
    else:
        raise ValueError(renorm)
    left, right, sk, s_all = split(A2, 'right' if going_right else 'left', max_bond, cutoff, gauge)
    return dict(A=A, Z=Z, psi=psi, S=S, dNLL=dNLL, A_new=A2, left=left, right=right, s=sk, s_all=s_all)


def sweep_schedule(L):
    """[0..L-2] + [L-2..0]  (TNutils.py:1597)."""
    return list(range(0, L - 1)) + list(range(L - 2, -1, -1))


def learning_epoch_cached(mps, imgs, epochs, lr, cache=None, batch_size=None, renorm='max',
                          max_bond=None, cutoff=1e-10, lr_update=None, update_wrap=None,
                          rescale='maxabs', rng=None, val_imgs=None, record=None, gauge=None):
    """TNutils.py:1570-1631.  Full-batch unless batch_size < N (then `rng.shuffle` plays the role of
    np.random.shuffle(guide), :1601-1602).  Mutates `mps` (a list) in place; returns (cost, lr).
    `record(step_index, k, going_right, out)` is called after every step if given."""
    d = mps[0].shape[-1]
    phys = embed(imgs, d)
    n = len(imgs)
    L = len(mps)
    if cache is None:
        cache = EnvCache(mps, phys, rescale)
    batch_size = n if batch_size is None else min(n, batch_size)
    guide = np.arange(n)
    cost = []
    step_no = 0
    # TNutils.py:1591 `lr = copy.copy(initial_lr)`: an mps_lr passed as lr is (shallow-)copied, so an update_wrap
    # closure that refers to the caller's object sees t == 0 forever (momentum never activates in the notebooks)
    import copy as _copy
    lr = _copy.copy(lr)
    for _ in range(epochs):
        going_right = True
        for k in sweep_schedule(L):
            if batch_size < n:
                (rng if rng is not None else np.random).shuffle(guide)
                mask = guide[:batch_size]
                lv, rv = cache.lenv[k][mask], cache.renv[k + 2][mask]
                pk, pk1 = phys[mask, k], phys[mask, k + 1]
            else:
                lv, rv = cache.lenv[k], cache.renv[k + 2]
                pk, pk1 = phys[:, k], phys[:, k + 1]
            cur_lr = getattr(lr, 'curr_lr', lr)
            out = two_site_step(mps, k, lv, rv, pk, pk1, cur_lr, going_right, renorm,
                                max_bond, cutoff, update_wrap, gauge=gauge)
            write_back(mps, k, out['left'], out['right'])
            if going_right:
                cache.advance_right(mps, k)
            else:
                cache.advance_left(mps, k + 1)
            if record is not None:
                record(step_no, k, going_right, out)
            step_no += 1
            if k == L - 2:
                going_right = False
        nll = compute_nll(mps, imgs if val_imgs is None else val_imgs, 0)
        if lr_update is not None:
            lr = lr_update(lr)
        cost.append(nll)
    return cost, lr


def compress(mps, max_bond):
    """TNutils.py:930-947: right-to-left two-site SVD sweep, absorb left, cap max_bond."""
    for k in range(len(mps) - 2, -1, -1):
        left, right, _, _ = split(merge(mps, k), 'left', max_bond)
        write_back(mps, k, left, right)
    return mps


class MpsLr:
    """`mps_lr` (TNutils.py:76-116): exponential annealing + the scalar-mean 'momentum' quirk."""

    def __init__(self, n_sites, lr0, momentum, s_factor):
        self.sites = n_sites
        self.lr0 = lr0
        self.curr_lr = lr0
        self.momentum = momentum
        self.t = 0
        self.s_factor = s_factor
        self.past_grad = (n_sites - 1) * [0]

    def new_epoch(self):
        self.t += 1
        self.curr_lr = self.lr0 * np.exp(-self.s_factor * self.t)

    def J(self, left_index, dNLL):
        J = self.momentum * self.past_grad[left_index] + dNLL if self.t > 0 else dNLL
        self.past_grad[left_index] = np.mean(dNLL)
        return J


# ----------------------------------------------------------------------------------------------
# samplers (TNutils.py:1829-1872 batched rule, :1874-1995 single rule, :2053-2116 reconstruct)
# ----------------------------------------------------------------------------------------------
def generate_samples(mps, U, rescale=False):
    """Batched ancestral sampler, TNutils.py:1829-1872, in float64.  `U` is (L,N): U[s] plays the
    role of the s-th `np.random.random(N)` call.  Pixel 1 at site s iff U[s,n] <= P(phys 0 | prefix).
    Assumes a right-canonical MPS (centre at 0).  `rescale=True` applies an exact power-of-two
    rescale of the running vectors (bit-identical decisions barring underflow)."""
    L = len(mps)
    N = U.shape[1]
    samples = np.zeros((N, L))
    m0 = site3(mps, 0)
    ampl = m0[0, :, 0]
    p = (ampl @ ampl) / np.sum(m0 * m0)
    acc = U[0] <= p
    samples[acc, 0] = 1
    half = np.where(acc[:, None], m0[0, :, 0][None, :], m0[0, :, 1][None, :])
    for s in range(1, L):
        m = site3(mps, s)
        ampl = half @ m[:, :, 0]
        cond = np.einsum('nb,nb->n', ampl, ampl) / np.einsum('nb,nb->n', half, half)
        acc = U[s] <= cond
        samples[acc, s] = 1
        if s == L - 1:
            return samples
        new = np.empty_like(ampl)
        new[acc] = ampl[acc]
        rej = ~acc
        if rej.any():
            new[rej] = half[rej] @ m[:, :, 1]
        if rescale:
            new, _ = _rescale(new, 'pow2')
        half = new
    return samples


def generate_sample(mps, u, start=None):
    """Single-image sampler, TNutils.py:1874-1995: tests pixel 0 (phys 1) first with a strict `<`.
    `u[j]` plays the role of the j-th np.random.rand().  Input must already be right-canonical
    (the reference calls mps.right_canonize() first, :1929).  `start` is an optional left boundary
    vector contracted into site 0 (used for known-prefix reconstruction)."""
    L = len(mps)
    m0 = site3(mps, 0)
    if start is not None:
        m0 = np.einsum('a,abp->bp', start, m0)[None]
    h0 = m0[0, :, 1]
    h1 = m0[0, :, 0]
    p0, p1 = h0 @ h0, h1 @ h1
    if L == 1:
        return [0] if u[0] < p0 / (p0 + p1) else [1]
    if u[0] < p0 / (p0 + p1):
        gen = [0]; prev = h0
    else:
        gen = [1]; prev = h1
    for k in range(1, L):
        m = site3(mps, k)
        new = prev @ m[:, :, 1]
        p = (new @ new) / (prev @ prev)
        if u[k] < p:
            gen.append(0)
        else:
            gen.append(1)
            new = prev @ m[:, :, 0]
        prev = new
    return gen


def reduced_chain(mps, img):
    """Contract every known pixel into its site and fuse known-only sites into the next unknown
    tensor, trailing known sites into the last unknown one (TNutils.py:2059-2109).  Returns a list of
    (Dl,Dr,d) tensors, one per unknown pixel."""
    d = mps[0].shape[-1]
    ph = embed(np.asarray(img), d)
    L = len(mps)
    chain = []
    carry = None   # product of known matrices waiting for the next unknown site
    for k in range(L):
        m = site3(mps, k)
        if ph[k] == UNKNOWN:
            t = m if carry is None else np.einsum('ab,bcp->acp', carry, m)
            chain.append(t); carry = None
        else:
            mk = m[:, :, ph[k]]
            carry = mk if carry is None else carry @ mk
    if carry is not None and chain:
        chain[-1] = np.einsum('abp,bc->acp', chain[-1], carry)
    return chain


def reconstruct(mps, img, u):
    """TNutils.py:2053-2116: normalise, reduce to the unknown-pixel chain, right-canonise it,
    run the single-image sampler, fill the -1 entries."""
    img = np.asarray(img)
    chain = reduced_chain(normalize_mps(mps), img)
    out = img.copy()
    if not chain:
        return out
    # the reduced chain may have non-trivial outer bonds; close them into plain vectors
    dl0 = chain[0].shape[0]
    dre = chain[-1].shape[1]
    if dl0 != 1:
        raise ValueError('reduced chain must start with a boundary of size 1')
    if dre != 1:
        raise ValueError('reduced chain must end with a boundary of size 1')
    red = right_canonize(from_site3(chain)) if len(chain) > 1 else from_site3(chain)
    if len(chain) == 1:
        t = chain[0]
        p0 = float(np.sum(t[:, :, 1] ** 2)); p1 = float(np.sum(t[:, :, 0] ** 2))
        gen = [0] if u[0] < p0 / (p0 + p1) else [1]
    else:
        gen = generate_sample(red, u)
    out[img == -1] = gen
    return out
