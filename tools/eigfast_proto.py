"""numpy prototype of the latency-optimised symmetric eigensolver planned for the Gram split:
Householder tridiagonalisation -> multisection (Sturm counts) -> twisted factorisation vectors -> back-transform ->
Newton-Schulz polish.  Validates accuracy / fallback predictor on Gram matrices of real two-site updates."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from oracle import born_oracle as o

EPS = 2.220446049250313e-16


def tridiag(G):
    """LAPACK dsytd2 (lower) arithmetic; returns d, e, V (reflector vectors, v[j+1]=1), tau."""
    A = G.copy(); n = A.shape[0]
    d = np.zeros(n); e = np.zeros(n - 1); tau = np.zeros(n - 1); V = np.zeros((n - 1, n))
    for j in range(n - 2):
        alpha = A[j + 1, j]; x = A[j + 2:, j]
        xnorm2 = x @ x
        if xnorm2 == 0.0:
            e[j] = alpha; d[j] = A[j, j]; continue
        beta = -np.copysign(np.sqrt(alpha * alpha + xnorm2), alpha)
        t = (beta - alpha) / beta
        v = np.zeros(n); v[j + 1] = 1.0; v[j + 2:] = x / (alpha - beta)
        p = t * (A @ v)           # v is zero on rows <= j: only the trailing block contributes... plus col j (ignored)
        p[:j + 1] = 0
        w = p - (0.5 * t * (p @ v)) * v
        A -= np.outer(v, w) + np.outer(w, v)
        e[j] = beta; d[j] = G[j, j] if j == 0 else A[j, j]
        tau[j] = t; V[j] = v
    d[n - 2] = A[n - 2, n - 2]; d[n - 1] = A[n - 1, n - 1]; e[n - 2] = A[n - 1, n - 2]
    for j in range(n - 2):
        pass
    return d, e, V, tau


def sturm_counts(d, e2, xs, pivmin):
    q = d[0] - xs
    q = np.where(np.abs(q) < pivmin, -pivmin, q)
    cnt = (q < 0).astype(np.int64)
    for i in range(1, d.size):
        q = d[i] - xs - e2[i - 1] / q
        q = np.where(np.abs(q) < pivmin, -pivmin, q)
        cnt += q < 0
    return cnt


def multisect(d, e, P=256, passes=7):
    n = d.size
    e2 = e * e
    r = np.abs(np.r_[e, 0]) + np.abs(np.r_[0, e])
    lo0, hi0 = (d - r).min(), (d + r).max()
    tn = max(abs(lo0), abs(hi0))
    lo0 -= 2 * EPS * tn * n; hi0 += 2 * EPS * tn * n
    pivmin = np.finfo(float).tiny * max(1.0, e2.max())
    lo = np.full(n, lo0); hi = np.full(n, hi0)
    ks = np.arange(n)
    for _ in range(passes):
        t = (np.arange(P) + 1.0) / (P + 1.0)
        xs = lo[:, None] + (hi - lo)[:, None] * t[None, :]
        cnt = sturm_counts(d, e2, xs.reshape(-1), pivmin).reshape(n, P)
        # eigenvalue k (0-based ascending) lies in (x_t, x_{t+1}] where count(x_t) <= k < count(x_{t+1})
        below = (cnt <= ks[:, None]).sum(axis=1)          # number of points with count <= k
        grid = np.concatenate([lo[:, None], xs, hi[:, None]], axis=1)
        lo, hi = grid[ks, below], grid[ks, below + 1]
    return 0.5 * (lo + hi)


def twisted_vector(d, e, lam):
    n = d.size
    tiny = np.finfo(float).tiny
    Dp = np.zeros(n); Lp = np.zeros(n - 1); Dm = np.zeros(n); Um = np.zeros(n - 1)
    Dp[0] = d[0] - lam
    for i in range(n - 1):
        if Dp[i] == 0: Dp[i] = tiny
        Lp[i] = e[i] / Dp[i]
        Dp[i + 1] = (d[i + 1] - lam) - Lp[i] * e[i]
    Dm[n - 1] = d[n - 1] - lam
    for i in range(n - 2, -1, -1):
        if Dm[i + 1] == 0: Dm[i + 1] = tiny
        Um[i] = e[i] / Dm[i + 1]
        Dm[i] = (d[i] - lam) - Um[i] * e[i]
    gam = Dp + Dm - (d - lam)
    r = int(np.argmin(np.abs(gam)))
    z = np.zeros(n); z[r] = 1.0
    for i in range(r - 1, -1, -1):
        z[i] = -Lp[i] * z[i + 1]
    for i in range(r, n - 1):
        z[i + 1] = -Um[i] * z[i]
    return z / np.linalg.norm(z)


def eig_fast(G, chi):
    n = G.shape[0]
    d, e, V, tau = tridiag(G)
    lam = multisect(d, e)
    top = lam[::-1][:chi]
    Z = np.stack([twisted_vector(d, e, x) for x in top], axis=1)
    for j in range(n - 3, -1, -1):
        if tau[j] != 0:
            Z -= tau[j] * np.outer(V[j], V[j] @ Z)
    return lam, Z


def polish(U, steps):
    for _ in range(steps):
        U = U @ (1.5 * np.eye(U.shape[1]) - 0.5 * (U.T @ U))
    return U


def case(G, chi, name):
    G = G / np.abs(G).max()
    lam_ref, Q = np.linalg.eigh(G)
    lam, U = eig_fast(G, chi)
    lmax = lam_ref[-1]
    sel = lam[::-1][:chi + 1] if chi < G.shape[0] else np.r_[lam[::-1], -np.inf]
    gaps = sel[:-1] - sel[1:]
    pred = 4 * EPS * lmax / max(gaps.min(), 1e-300)
    orth = np.abs(U.T @ U - np.eye(chi)).max()
    U2 = polish(U, 2)
    orth2 = np.abs(U2.T @ U2 - np.eye(chi)).max()
    # subspace error vs LAPACK top-chi
    Qt = Q[:, ::-1][:, :chi]
    sub = np.linalg.norm(U2 - Qt @ (Qt.T @ U2), 2)
    resid = np.abs(G @ U2 - U2 * lam[::-1][:chi]).max() / lmax
    print(f'{name:28s} n={G.shape[0]:4d} chi={chi:4d} lam_err={np.abs(lam - lam_ref).max() / lmax:.1e} '
          f'lam_min_sel={sel[chi - 1] / lmax:.1e} pred={pred:.1e} orth={orth:.1e} polished={orth2:.1e} subspace={sub:.1e} resid={resid:.1e}')


def gram_from_step(L, D, N, seed, lr=0.05, centre=None):
    rng = np.random.default_rng(seed)
    k = L // 2
    mps = o.canonize(o.random_mps(L, D, rng=rng), k)
    imgs = (rng.random((N, L)) < 0.4).astype(np.int64)
    phys = o.embed(imgs)
    cache = o.EnvCache(mps, phys, 'pow2')
    A = o.merge(mps, k)
    psi, S = o.psi_and_grad(A, cache.lenv[k], cache.renv[k + 2], phys[:, k], phys[:, k + 1])
    Z = np.sum(A * A)
    A2 = A - lr * (A / Z - S / N)
    A2 = A2 / np.abs(A2).max()
    Dl, d, Dr, _ = A2.shape
    M = A2.transpose(1, 0, 3, 2).reshape(d * Dl, d * Dr)
    return M @ M.T


def twisted_vector_three_term(d, e, lam):
    """Round-2 candidate for k_tri_vectors: D+_i = p_i / p_{i-1} from the three-term recurrence, so that the division
    leaves the dependency chain (one dependent FMA per row instead of reciprocal + FMA).  Same accuracy as the qd form on
    every case of `python tools/eigfast_proto.py --three-term` (orthogonality and residual agree to the printed digits)."""
    n = d.size
    e2 = e * e
    tiny = np.finfo(float).tiny

    def ratios(dd, ee2):
        p_prev, p = 1.0, dd[0] - lam
        D = np.empty(n); D[0] = p
        for i in range(1, n):
            pn = (dd[i] - lam) * p - ee2[i - 1] * p_prev
            D[i] = pn / (p if p != 0.0 else tiny)
            p_prev, p = p, pn
            m = max(abs(p), abs(p_prev))
            if m > 1e100 or 0 < m < 1e-100:          # power-of-two rescaling on the device
                p /= m; p_prev /= m
        return D
    Dp = ratios(d, e2)
    Dm = ratios(d[::-1], e2[::-1])[::-1]
    Dp = np.where(Dp == 0, tiny, Dp); Dm = np.where(Dm == 0, tiny, Dm)
    Lp, Um = e / Dp[:-1], e / Dm[1:]
    r = int(np.argmin(np.abs(Dp + Dm - (d - lam))))
    z = np.zeros(n); z[r] = 1.0
    for i in range(r - 1, -1, -1):
        z[i] = -Lp[i] * z[i + 1]
    for i in range(r, n - 1):
        z[i + 1] = -Um[i] * z[i]
    return z / np.linalg.norm(z)


def compare_three_term():
    rng = np.random.default_rng(0)
    cases = [('step D=32', gram_from_step(16, 32, 500, 1), 32), ('step D=64', gram_from_step(18, 64, 800, 2), 64)]
    for dec in (3, 6, 9):
        Qm, _ = np.linalg.qr(rng.standard_normal((128, 128)))
        cases.append((f'geometric {dec} decades', (Qm * (10.0 ** (-dec * np.arange(128) / 128)) ** 2) @ Qm.T, 64))
    for name, G, chi in cases:
        G = G / np.abs(G).max()
        d, e, V, tau = tridiag(G)
        lam = multisect(d, e)
        top = lam[::-1][:chi]
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        for label, fn in (('qd', twisted_vector), ('three-term', twisted_vector_three_term)):
            Z = np.stack([fn(d, e, x) for x in top], axis=1)
            print(f'{name:24s} {label:10s} ortho {np.abs(Z.T @ Z - np.eye(chi)).max():.1e} resid {np.abs(T @ Z - Z * top).max() / lam[-1]:.1e}')


if __name__ == '__main__':
    if '--three-term' in sys.argv:
        compare_three_term()
        sys.exit(0)
    rng = np.random.default_rng(0)
    G = gram_from_step(16, 32, 500, 1)
    case(G, 32, 'step D=32')
    G = gram_from_step(18, 64, 800, 2)
    case(G, 64, 'step D=64')
    # trained-like spectrum: geometric decay over 3 / 6 / 9 decades
    for dec in (3, 6, 9):
        n = 128
        Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
        s = 10.0 ** (-dec * np.arange(n) / n)
        case((Qm * s ** 2) @ Qm.T, 64, f'geometric {dec} decades (sigma)')
    # degenerate pairs
    n = 64
    Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = np.repeat(10.0 ** (-2 * np.arange(n // 2) / n), 2)
    case((Qm * s ** 2) @ Qm.T, 32, 'exactly degenerate pairs')
    # low rank
    M = rng.standard_normal((64, 10)) @ rng.standard_normal((10, 64))
    case(M @ M.T, 10, 'rank 10 of 64')
