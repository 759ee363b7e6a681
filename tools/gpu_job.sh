python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.log; tail -1 gpurun_out/bench13.log | cut -c1-3000
