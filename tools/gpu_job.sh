timeout 900 python -m pytest tests/test_stream_gpu.py tests/test_inflate_gpu.py -m gpu -x -q 2>&1 | tail -4
python tools/probe_stream_rate.py 2>&1 | tail -2
