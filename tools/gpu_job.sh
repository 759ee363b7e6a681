python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench16.json 2> gpurun_out/bench16.log; tail -1 gpurun_out/bench16.log | cut -c1-1500
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
