"""GPU check of the on-chip eigensolver (bm_eig_sym) against numpy.linalg.eigh + timing against cuSOLVER syevd."""
import json, sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import mps_born_machines_b200 as bm


def gram(n, q, rng, decades=None):
    if decades is None:
        Y = rng.standard_normal((q, n))
    else:
        Q1, _ = np.linalg.qr(rng.standard_normal((q, n)))
        Q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
        Y = (Q1 * 10.0 ** (-decades * np.arange(n) / n)) @ Q2.T
    return Y.T @ Y


def main():
    rng = np.random.default_rng(0)
    eng = bm.BornEngine(4, 256)
    out = []
    cases = [('n2', gram(2, 5, rng), 2), ('n3', gram(3, 5, rng), 3), ('n12', gram(12, 20, rng), 6), ('n33', gram(33, 40, rng), 33),
             ('n64', gram(64, 64, rng), 32), ('n200', gram(200, 200, rng), 100), ('n511', gram(511, 600, rng), 255),
             ('n512', gram(512, 512, rng), 256), ('n512_3dec', gram(512, 512, rng, 3), 256), ('n512_6dec', gram(512, 512, rng, 6), 256),
             ('diag', np.diag(np.arange(1.0, 41.0)), 40), ('blockdiag', np.kron(np.eye(2), gram(16, 20, rng)) + np.diag(np.r_[np.zeros(16), np.ones(16) * 0.37]), 32),
             ('rank5of48', (lambda M: M @ M.T)(rng.standard_normal((48, 5))), 5)]
    for name, G, m in cases:
        n = G.shape[0]
        lam_ref, Q = np.linalg.eigh(G)
        lam, U, info = eng.eig_sym(G, m, on_chip=True)
        lam2, U2, info2 = eng.eig_sym(G, m, on_chip=False)
        runs = [eng.eig_sym(G, m, on_chip=True)[2] for _ in range(3)]
        ms = min(r['ms'] for r in runs)
        ms2 = min(eng.eig_sym(G, m, on_chip=False)[2]['ms'] for _ in range(3))
        lmax = abs(lam_ref).max()
        top = lam[::-1][:m]
        rec = {'case': name, 'n': n, 'm': m,
               'lam_err': float(np.abs(lam - lam_ref).max() / lmax),
               'ortho_before_polish': info['ortho_dev'],
               'ortho': float(np.abs(U.T @ U - np.eye(m)).max()),
               'resid': float(np.abs(G @ U - U * top).max() / lmax),
               'syevd_resid': float(np.abs(G @ U2 - U2 * lam2[::-1][:m]).max() / lmax),
               'ms_on_chip': ms, 'ms_syevd': ms2, 'stage_ms': {k: round(v, 4) for k, v in runs[-1]['stage_ms'].items()}}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    eng.close()


if __name__ == '__main__':
    main()
