python bench.py --steps 1 --warmup 1 --no-extra --no-verify > gpurun_out/plain.json 2> gpurun_out/plain.log && ncu --set full --clock-control none --import-source on -k regex:"k_lz_link|k_lz_walk|k_huff_build|k_huff_pack" -s 8 -c 4 -o gpurun_out/r1_f python bench.py --steps 1 --warmup 1 --no-extra --no-verify > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_f.csv python bench.py --steps 2 --warmup 1 --no-extra --no-verify > gpurun_out/ncu_l4.log 2>&1
wc -l gpurun_out/r1_launches_f.csv
