"""Development A/B (GPU): the flat lazy walk against the nested one -- identical output bytes (sha256 per level and
corpus) and the time per level.  Each variant in its own process (the knob is read once)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, hashlib
sys.path.insert(0, %r)
import torch
from zlib_b200 import load, binding as zb, synth
L = load(); assert L.dll.zb200_init(0) == 0
s = torch.cuda.current_stream()
out = {}
for name, kind, n, seed in (("mixed128m", 1, 128 << 20, 1), ("text64m", 0, 64 << 20, 7), ("noise8m", 2, 8 << 20, 3)):
    d = torch.from_numpy(synth.synth(n, kind, seed)).cuda()
    cap = L.compress_bound(n) + 64
    o = torch.empty(cap, dtype=torch.uint8, device="cuda")
    for lv in (4, 5, 6, 7, 9):
        if lv == 9 and n > (64 << 20): continue
        clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, lv, zb.WRAP_ZLIB, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, lv, zb.WRAP_ZLIB, s)
        e1.record(s); torch.cuda.synchronize()
        h = hashlib.sha256(o[:clen].cpu().numpy().tobytes()).hexdigest()[:12]
        out[f"{name}_L{lv}"] = [round(e0.elapsed_time(e1), 2), int(clen), h]
    del d, o
print(json.dumps(out))
''' % ROOT
res = {}
for flat in ("0", "1"):
    env = dict(os.environ, ZB200_LAZY_FLAT=flat)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=900)
    try:
        res[flat] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print("variant", flat, "failed:", r.stdout[-300:], r.stderr[-800:]); sys.exit(1)
same = True
for k in res["0"]:
    a, b = res["0"][k], res["1"][k]
    ok = a[1:] == b[1:]
    same &= ok
    print(f"{k:16s} nested {a[0]:8.2f} ms  flat {b[0]:8.2f} ms  x{a[0] / b[0]:.2f}  bytes {a[1]} {'identical' if ok else 'DIFFERENT ' + str(b[1])}")
print("ALL IDENTICAL" if same else "MISMATCH")
