python tools/probe_single.py 256 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_inflate_segments|k_resolve_tails|k_find_blocks" -c 3 -o gpurun_out/r1_single python tools/probe_single.py 256 > gpurun_out/ncu_single.log 2>&1
tail -2 gpurun_out/ncu_single.log
