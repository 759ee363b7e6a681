timeout 600 python tools/probe_single_ref.py 2>&1 | tail -12
