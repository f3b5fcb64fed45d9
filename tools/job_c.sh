mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_pytest_c.log; cat gpurun_out/r02_pytest_c.log
BM_EIG_DEBUG=1 python tools/eig_one.py 2>&1 | grep -E "sytrd phases|sub-phases|stage_ms" | tail -3 | tee gpurun_out/r02_sytrd_phases.log
for v in fused bulk; do
BM_SYTRD=$v python tools/test_eig.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l)
    if r['case'] in ('n512','n200','n64'): print('$v', r['case'], 'ms %.3f'%r['ms_on_chip'], r['stage_ms'])
"
done
