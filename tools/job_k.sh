BM_EIG_DEBUG=2 timeout 100 python tools/eig_one.py 2>&1 | grep -E "phases|phase C" | tail -4
timeout 100 python tools/test_eig.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l)
    if r['case'] in ('n512','n200','n64','n33'): print(r['case'], 'lam %.1e ortho %.1e resid %.1e'%(r['lam_err'],r['ortho'],r['resid']), 'ms %.3f'%r['ms_on_chip'], r['stage_ms'])
"
