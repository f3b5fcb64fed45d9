for tau in 1e-7 1e-8; do
BM_GRAM_TAU=$tau timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu --no-others --no-sampler --no-sweep 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tau', r['ms_per_step'], sum(1 for x in r['gram_levels'] if x>1), 'of', len(r['gram_levels']), r['split_eig'])"
done
