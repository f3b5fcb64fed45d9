"""Host-side MPS helpers of the facade: initialisation and canonicalisation (a2: `initialize_mps`,
TNutils.py:386-424 -> quimb MPS_rand_state + canonize).  One-off set-up work done with numpy QR on the host;
the per-step hot path never comes through here."""
import numpy as np


def site3(mps, k):
    t = mps[k]
    if t.ndim == 3:
        return t
    return t.reshape(1, t.shape[0], t.shape[1]) if k == 0 else t.reshape(t.shape[0], 1, t.shape[1])


def squeeze_chain(mats):
    L = len(mats)
    out = []
    for k, t in enumerate(mats):
        if k == 0:
            out.append(np.ascontiguousarray(t[0]))
        elif k == L - 1:
            out.append(np.ascontiguousarray(t[:, 0, :]))
        else:
            out.append(np.ascontiguousarray(t))
    return out


def canonize(mps, centre):
    """QR sweeps: sites < centre left isometries, sites > centre right isometries (quimb canonize)."""
    L = len(mps)
    m = [site3(mps, k).copy() for k in range(L)]
    for k in range(0, centre):
        dl, dr, d = m[k].shape
        q, r = np.linalg.qr(m[k].transpose(0, 2, 1).reshape(dl * d, dr))
        m[k] = q.reshape(dl, d, -1).transpose(0, 2, 1)
        m[k + 1] = np.tensordot(r, m[k + 1], axes=([1], [0]))
    for k in range(L - 1, centre, -1):
        dl, dr, d = m[k].shape
        q, r = np.linalg.qr(m[k].reshape(dl, dr * d).T)
        m[k] = q.T.reshape(-1, dr, d)
        m[k - 1] = np.tensordot(m[k - 1], r.T, axes=([1], [0])).transpose(0, 2, 1)
    return squeeze_chain(m)


def random_mps(L, bond_dim, d=2, rng=None, centre=0, decades=0.0):
    """i.i.d. normal site tensors with every bond = bond_dim (like qtn.MPS_rand_state), canonised to
    `centre` and normalised (<mps|mps> = |centre tensor|^2 = 1).  decades > 0: the right bond index b of every site is
    scaled by 10^(-decades * b / bond_dim) first, which gives every cut a Schmidt spectrum decaying over about that many
    decades (a stand-in for a trained model; a plain random MPS has an almost flat spectrum)."""
    rng = np.random.default_rng(0) if rng is None else rng
    mats = [rng.standard_normal((1 if k == 0 else bond_dim, 1 if k == L - 1 else bond_dim, d)) / np.sqrt(bond_dim * d)
            for k in range(L)]
    if decades > 0:
        for k in range(L - 1):
            w = 10.0 ** (-decades * np.arange(mats[k].shape[1]) / bond_dim)
            w = w / np.sqrt(np.mean(w * w))                    # unit RMS: the norm of the chain stays O(1) per site
            mats[k] = mats[k] * w[None, :, None]
    mps = squeeze_chain(mats)
    if centre is not None and 0 <= centre < L:
        mps = canonize(mps, centre)
        mps[centre] = mps[centre] / np.linalg.norm(mps[centre])
    return mps


def synthetic_images(n, L=784, seed=1234, n_proto=64, p_on=0.15, p_flip=0.05):
    """Synthetic binarised 28x28-like images (no MNIST offline): 64 prototypes with 15% ink, 5% bit flips."""
    rng = np.random.default_rng(seed)
    protos = rng.random((n_proto, L)) < p_on
    c = rng.integers(n_proto, size=n)
    flips = rng.random((n, L)) < p_flip
    return (protos[c] ^ flips).astype(np.uint8)
