python -m pytest tests/test_deflate_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -3
DEV_NO_EDGE=1 timeout 300 python tools/dev_check.py 1024 1 2>&1 | tail -1
