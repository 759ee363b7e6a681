python tools/probe_codec.py 1024 2>&1 | head -1
ZB200_SLAB_LANES=3 python tools/probe_codec.py 1024 2>&1 | head -1
ZB200_SLAB_LANES=3 python -m pytest tests/test_deflate_gpu.py -m gpu -x -q 2>&1 | tail -2
