python -m pytest tests/test_deflate_gpu.py -m gpu -x -q 2>&1 | tail -3
for k in 1 0; do DEV_KIND=$k DEV_NO_EDGE=1 timeout 300 python tools/dev_check.py 1024 1,6 2>&1 | tail -2; done
