python -m pytest tests/test_inflate_gpu.py tests/test_stream_gpu.py tests/test_abi.py -x -q 2>&1 | tail -2
python tools/probe_inflate_host.py 32768 2>&1 | tail -4
