timeout 900 python -m pytest tests/test_inflate_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/probe_inflate_host.py 32768 2>&1 | tail -5
