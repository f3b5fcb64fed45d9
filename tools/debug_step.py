import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import born_oracle as o
import mps_born_machines_b200 as bm
L, N = 8, 40
rng = np.random.default_rng(31)
mps = o.canonize(o.random_mps(L, 2, rng=rng), 0)
imgs = (rng.random((N, L)) < 0.4).astype(np.int64)
phys = o.embed(imgs)
eng = bm.BornEngine(L, 16); eng.set_mps(mps); eng.set_data(imgs)
mps_o = [t.copy() for t in mps]
cache = o.EnvCache(mps_o, phys, 'pow2')
gr = True
for k in o.sweep_schedule(L)[:4]:
    ref = o.two_site_step(mps_o, k, cache.lenv[k], cache.renv[k+2], phys[:,k], phys[:,k+1], 0.3, gr, 'max', max_bond=16)
    S = eng.step_grad(k)
    A = eng.get_merged_at(k)
    print('k', k, 'A err', np.abs(A-ref['A']).max(), 'S err', np.abs(eng.grad_to_ref(k,S)-ref['S']).max(), np.abs(ref['S']).max())
    out = eng.step_update(k, S, N, 0.3, gr, 'max', 16, 1e-10, 0.0, True)
    print('   Z', out['Z'], ref['Z'], 'c', out['renorm_div'], (ref['A']-0.3*ref['dNLL']).max(), 'chi', out['chi'], ref['s'].size)
    print('   s', out['s_all'], ref['s_all'])
    o.write_back(mps_o, k, ref['left'], ref['right'])
    prod_ref = np.einsum('abp,bcq->apcq', o.site3(mps_o,k), o.site3(mps_o,k+1))
    prod = np.einsum('abp,bcq->apcq', eng.get_site(k,False), eng.get_site(k+1,False))
    print('   prod err', np.abs(prod-prod_ref).max(), 'shapes', eng.get_site(k,False).shape, eng.get_site(k+1,False).shape)
    if gr: cache.advance_right(mps_o, k)
    else: cache.advance_left(mps_o, k+1)
    eng.env_advance(k, gr)
