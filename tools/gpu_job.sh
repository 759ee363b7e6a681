python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.log; echo bench rc=$?
tail -1 gpurun_out/bench7.log | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r1_launches_c.csv python bench.py --steps 1 --warmup 1 --no-extra --no-verify > gpurun_out/ncu_l.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:'k_lz_|k_huff' --launch-skip 8 -c 4 -o gpurun_out/r1_e python bench.py --steps 1 --warmup 1 --no-extra --no-verify > gpurun_out/ncu_e.log 2>&1; echo rc=$?
