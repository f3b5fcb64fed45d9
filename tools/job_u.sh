timeout 200 python -m pytest tests/test_gpu_facade.py -x -q -m gpu -k "uploads_only" 2>&1 | grep -E "Error|error|assert|^E " | head -20
