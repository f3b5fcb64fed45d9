timeout 100 python tools/test_eig.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l)
    print(r['case'], 'lam %.1e ortho %.1e resid %.1e'%(r['lam_err'],r['ortho'],r['resid']), 'ms %.3f'%r['ms_on_chip'], r['stage_ms']['sytrd'])
"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "eigensolver or split" 2>&1 | tail -3
