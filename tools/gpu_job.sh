python bench.py > gpurun_out/bench15.json 2> gpurun_out/bench15.log; tail -1 gpurun_out/bench15.log | cut -c1-1200
