python -m pytest tests/test_checksum_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/probe_checksum.py 2>&1 | tail -5
