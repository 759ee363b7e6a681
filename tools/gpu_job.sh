python bench.py > gpurun_out/bench9.json 2> gpurun_out/bench9.log; echo bench rc=$?
tail -2 gpurun_out/bench9.log | cut -c1-700
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
