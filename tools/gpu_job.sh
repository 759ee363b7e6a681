timeout 900 python -m pytest tests/test_deflate_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/probe_batch.py 10000 8 2>&1 | grep "pageable" | tail -3
