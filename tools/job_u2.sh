timeout 300 python -m pytest tests/test_gpu_facade.py -x -q -m gpu 2>&1 | tail -3
