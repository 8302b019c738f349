set -x
timeout 600 python -m pytest tests/test_gpu_masked.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
