python bench.py --steps 2 --warmup 1 --no-extra --no-verify > gpurun_out/plain.json 2> gpurun_out/plain.log && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_e.csv python bench.py --steps 2 --warmup 1 --no-extra --no-verify > gpurun_out/ncu_l3.log 2>&1
wc -l gpurun_out/r1_launches_e.csv
