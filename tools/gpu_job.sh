python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.log; echo bench rc=$?
tail -4 gpurun_out/bench4.log | cut -c1-400
