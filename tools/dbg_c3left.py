"""per-step singular-value agreement of the engine against the oracle at the c3 shape, both directions"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import born_oracle as o
import mps_born_machines_b200 as bm

def run(going_right, svd='gram', L=22, D=256, N=8192, seed=301, lr=1e-3):
    bonds = [9, 10, 11] if going_right else [11, 10, 9]
    centre = bonds[0] if going_right else bonds[0] + 1
    rng = np.random.default_rng(seed)
    mps = o.canonize(o.random_mps(L, D, d=2, rng=rng), centre)
    imgs = o.synthetic_images(N, L)
    phys = o.embed(imgs, 2)
    eng = bm.BornEngine(L, D, phys_dim=2, svd=svd)
    eng.set_mps(mps); eng.set_data(imgs)
    cache = o.EnvCache(mps, phys, 'pow2')
    mps_o = [t.copy() for t in mps]
    for k in bonds:
        ref = o.two_site_step(mps_o, k, cache.lenv[k], cache.renv[k + 2], phys[:, k], phys[:, k + 1], lr, going_right, 'l2', max_bond=D)
        o.write_back(mps_o, k, ref['left'], ref['right'])
        (cache.advance_right(mps_o, k) if going_right else cache.advance_left(mps_o, k + 1))
        out = eng.step(k, lr, going_right, 'l2', max_bond=D, want_s=True)
        rel = np.abs(out['s'] - ref['s']) / ref['s']
        sl = (np.log(np.abs(ref['psi'])) + cache.ls_l[k] + cache.ls_r[k + 2]).sum()
        print(svd, 'right' if going_right else 'left', k, 'chi', out['chi'], ref['s'].size, 'max rel s %.2e' % rel.max(), 'argmax', int(rel.argmax()),
              'Z %.3e' % abs(out['Z'] / ref['Z'] - 1), 'sumlog %.3e' % abs(out['sum_logpsi'] / sl - 1), 'minpsi %.2e' % np.abs(ref['psi']).min(),
              's range %.3e %.3e' % (ref['s'][0], ref['s'][-1]), 'levels', eng.eig_stats() and eng.last_gram_levels, flush=True)
    eng.close()

for svd in ('gram', 'gram_syevd', 'gesvdj'):
    for gr in (True, False):
        run(gr, svd)
