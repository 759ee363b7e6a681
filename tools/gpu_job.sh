python -m pytest tests/test_deflate_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/probe_codec.py 1024 2>&1 | head -2
