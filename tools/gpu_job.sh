DEV_NO_EDGE=1 timeout 300 python tools/dev_check.py 1024 1,6 2>&1 | tail -2
