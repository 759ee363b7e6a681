python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench17.json 2> gpurun_out/bench17.log; grep -E "deflate L1|e2e compress2" gpurun_out/bench17.log; python -c "
import json; d=json.load(open('gpurun_out/bench17.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernels_ms'])"
