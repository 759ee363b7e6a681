python -m pytest tests/test_checksum_gpu.py tests/test_inflate_gpu.py tests/test_deflate_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/probe_batch.py 10000 8 2>&1 | tail -3
