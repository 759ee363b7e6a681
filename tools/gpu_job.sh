python -m pytest tests/test_deflate_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-extra > gpurun_out/bench8.json 2> gpurun_out/bench8.log; echo bench rc=$?
grep -E "deflate L1|e2e|verified" gpurun_out/bench8.log
