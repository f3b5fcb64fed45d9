for lib in libbm.so libbm_t128.so; do
BM_LIB_PATH=$PWD/mps_born_machines_b200/_lib/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-others --no-sampler --no-sweep 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=r['roofline']['kernels']; print('$lib', r['ms_per_step'], {n:(round(k[n]['ms'],4), round(k[n].get('tflops',0),1)) for n in ('psi','env_advance','grad')})"
done
