timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/probe_single.py 1024 2>&1 | grep -v "^{" | tail -6
