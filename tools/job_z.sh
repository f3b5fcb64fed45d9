timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "level_threshold or small_singular or split" 2>&1 | tail -12
for t2 in 1e-11 1; do BM_GRAM_TAU2=$t2 timeout 200 python tools/tau_experiment.py 2>&1 | tail -1 | cut -c 1-420; done
