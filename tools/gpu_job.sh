python -m pytest tests/test_checksum_gpu.py tests/test_deflate_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/probe_checksum.py 2>&1 | tail -6
