"""Development sweep (GPU): level-1 search budget and walk-kernel variants -> ms per GiB and output size.
Every configuration runs in its own process because the knobs are read once (env)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, ctypes as C
sys.path.insert(0, %r)
import torch
from zlib_b200 import load, binding as zb, synth
L = load(); assert L.dll.zb200_init(0) == 0
s = torch.cuda.current_stream()
out = {}
for name, kind, n, seed in (("mixed1g", 1, 1 << 30, 1), ("text64m", 0, 64 << 20, 7)):
    d = torch.from_numpy(synth.synth(n, kind, seed)).cuda()
    cap = L.compress_bound(n) + 64
    o = torch.empty(cap, dtype=torch.uint8, device="cuda")
    for lv in (1,):
        for _ in range(2): clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, lv, zb.WRAP_ZLIB, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, lv, zb.WRAP_ZLIB, s)
        e1.record(s); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        L.profile(True); L.deflate(d.data_ptr(), n, o.data_ptr(), cap, lv, zb.WRAP_ZLIB, s); pr = L.profile_report(); L.profile(False)
        walk = sum(v[0] for k, v in pr.items() if "walk" in k)
        out[name] = {"ms": round(ms, 3), "GBps": round(n / ms / 1e6, 2), "bytes": int(clen), "walk_ms": round(walk, 3)}
    del d, o
print(json.dumps(out))
''' % ROOT
configs = [{}]
if len(sys.argv) > 1 and sys.argv[1] == "link":
    configs += [{"ZB200_LINK_SEG_CHUNKS": "1"}, {"ZB200_LINK_SEG_CHUNKS": "4"}]
else:
    configs += [{"ZB200_WALK_PERSIST": "0"}]
    for chain in (1, 3, 4):
        configs.append({"ZB200_L1_CHAIN": str(chain)})
    for nice in (6, 16, 32):
        configs.append({"ZB200_L1_NICE": str(nice)})
for cfg in configs:
    env = dict(os.environ); env.update(cfg)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(json.dumps(cfg), line, flush=True)
