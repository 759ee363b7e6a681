python bench.py > gpurun_out/bench12.json 2> gpurun_out/bench12.log; tail -1 gpurun_out/bench12.log | cut -c1-2500
