set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.log; echo bench rc=$?
tail -3 gpurun_out/bench2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-verify > gpurun_out/ncu_l.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:'k_lz_|k_huff|k_inflate' -c 8 -o gpurun_out/r1_codec python tools/probe_codec.py 64 > gpurun_out/ncu_f.log 2>&1; echo rc=$?
ls -la gpurun_out
