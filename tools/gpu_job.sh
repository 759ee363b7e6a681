python bench.py > gpurun_out/bench20.json 2> gpurun_out/bench20.log; python -c "
import json; d=json.load(open('gpurun_out/bench20.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']); print({k:(v.get('GBps') if isinstance(v,dict) else v) for k,v in d['extra'].items()})"
