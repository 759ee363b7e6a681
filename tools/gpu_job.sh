timeout 900 python -m pytest tests/test_stream_gpu.py -m gpu -x -q -k "ten_thousand" --durations=3 2>&1 | tail -8
