for tau in 1e-7 2.5e-8 1e-8; do BM_GRAM_TAU=$tau timeout 200 python tools/tau_experiment.py 2>&1 | tail -1; done
