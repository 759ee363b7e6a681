timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/probe_pageable.py 2>&1 | grep "pageable" | tail -5
