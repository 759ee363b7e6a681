python -m pytest tests/test_deflate_gpu.py -m gpu -x -q -k concurrent 2>&1 | tail -8
