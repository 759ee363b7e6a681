python bench.py > gpurun_out/bench11.json 2> gpurun_out/bench11.log; tail -3 gpurun_out/bench11.log | cut -c1-1500
