"""Development sweep (GPU): [level, {env}] configurations -> ms and output bytes on 256 MiB mixed and 64 MiB text (each
configuration in its own process; sizes of the reference build at every level are in profiles/r2_sweeps.md)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import torch
from zlib_b200 import load, binding as zb, synth
L = load(); assert L.dll.zb200_init(0) == 0
level = int(sys.argv[1])
s = torch.cuda.current_stream()
out = {}
for name, kind, n, seed in (("mixed256m", 1, 256 << 20, 1), ("text64m", 0, 64 << 20, 7)):
    d = torch.from_numpy(synth.synth(n, kind, seed)).cuda()
    cap = L.compress_bound(n) + 64
    o = torch.empty(cap, dtype=torch.uint8, device="cuda")
    clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, level, zb.WRAP_ZLIB, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(3): clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, level, zb.WRAP_ZLIB, s)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out[name] = [round(ms, 3), round(n / ms / 1e6, 1), int(clen)]
    del d, o
print(json.dumps(out))
''' % ROOT
arg = sys.argv[1] if len(sys.argv) > 1 else json.dumps([[l, {}] for l in range(1, 10)])   # a JSON list, @file, or every level at its defaults
for level, cfg in json.load(open(arg[1:])) if arg.startswith("@") else json.loads(arg):
    env = dict(os.environ); env.update({k: str(v) for k, v in cfg.items()})
    r = subprocess.run([sys.executable, "-c", CHILD, str(level)], env=env, capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(level, json.dumps(cfg), line, flush=True)
