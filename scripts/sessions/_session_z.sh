timeout 900 python -m pytest tests/test_gpu_statistics.py -m gpu -q -s -x -k "reference_distributions" 2>&1 | tail -8
