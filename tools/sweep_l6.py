"""Development sweep (GPU): level-6 chain budget -> ms per GiB and output size (each configuration in its own process)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import torch
from zlib_b200 import load, binding as zb, synth
L = load(); assert L.dll.zb200_init(0) == 0
s = torch.cuda.current_stream()
out = {}
for name, kind, n, seed in (("mixed256m", 1, 256 << 20, 1), ("text64m", 0, 64 << 20, 7)):
    d = torch.from_numpy(synth.synth(n, kind, seed)).cuda()
    cap = L.compress_bound(n) + 64
    o = torch.empty(cap, dtype=torch.uint8, device="cuda")
    clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, 6, zb.WRAP_ZLIB, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(3): clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, 6, zb.WRAP_ZLIB, s)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out[name] = {"ms": round(ms, 3), "GBps": round(n / ms / 1e6, 2), "bytes": int(clen)}
    del d, o
print(json.dumps(out))
''' % ROOT
for cfg in ({}, {"ZB200_LAZY_GLOBAL_MAX": "64"}, {"ZB200_L6_CHAIN": "24"}, {"ZB200_L6_CHAIN": "24", "ZB200_LAZY_GLOBAL_MAX": "64"}, {"ZB200_L6_CHAIN": "16"}):
    env = dict(os.environ); env.update(cfg)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(json.dumps(cfg), line, flush=True)
