"""Would a TF32 environment chain (tcgen05 kind::tf32 for sequential_update / psi) meet the 1e-3 parity class?
Emulates the arithmetic on the CPU: operands truncated to TF32 (10-bit mantissa, as the tensor core reads FP32 bits),
products accumulated in FP32, environments stored in FP32, power-of-two rescaling as in the engine.  Reports the error of
ln|psi| per sample, of the NLL, and of the two-site gradient direction against the FP64 oracle on a 784-site chain."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import born_oracle as o


ROUND = os.environ.get('TF32_ROUND', 'trunc')       # 'trunc': what kind::tf32 does with raw FP32 bits; 'rna': cvt.rna.tf32.f32 first


def tf32(x):
    a = np.asarray(x, dtype=np.float32).copy()
    v = a.view(np.uint32)
    if ROUND == 'rna':
        v += np.uint32(0x1000)                       # round half away from zero on the 13 dropped bits
    v &= np.uint32(0xFFFFE000)
    return a


def chain_tf32(mps, phys, upto):
    """left environments up to bond `upto` with TF32 products / FP32 storage; returns (env, log-scale)"""
    N = phys.shape[0]
    env = np.ones((N, 1), dtype=np.float32)
    ls = np.zeros(N)
    for j in range(upto):
        m = o.site3(mps, j)
        out = np.empty((N, m.shape[1]), dtype=np.float32)
        for x in range(m.shape[2]):
            sel = phys[:, j] == x
            if sel.any():
                out[sel] = (tf32(env[sel]).astype(np.float32) @ tf32(m[:, :, x]).astype(np.float32)).astype(np.float32)
        e = np.frexp(np.abs(out).max(axis=1))[1]
        out = np.ldexp(out, -e[:, None]).astype(np.float32)
        ls += e * np.log(2.0)
        env = out
    return env, ls


def main():
    L, D, N = 784, int(os.environ.get('D', 48)), 256
    rng = np.random.default_rng(5)
    mps = o.canonize(o.random_mps(L, D, rng=rng), L - 1)
    imgs = o.synthetic_images(N, L)
    phys = o.embed(imgs)
    _, lp = o.logpsi(mps, phys)                       # FP64 ln|psi|
    env, ls = chain_tf32(mps, phys, L)
    lp32 = np.log(np.abs(env[:, 0].astype(np.float64))) + ls
    err = np.abs(lp32 - lp)
    res = {'L': L, 'D': D, 'N': N, 'operand_rounding': ROUND,
           'ln_psi_abs_err_max': float(err.max()), 'ln_psi_abs_err_mean': float(err.mean()),
           'ln_psi_rel_err_max': float((err / np.abs(lp)).max()),
           'nll_fp64': float(-2 * lp.mean()), 'nll_tf32_chain': float(-2 * lp32.mean()),
           'nll_rel_err': float(abs(lp32.mean() - lp.mean()) / abs(lp.mean())),
           'psi_rel_err_max': float(np.abs(np.expm1(lp32 - lp)).max())}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
