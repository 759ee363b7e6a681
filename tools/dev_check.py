"""Development probe (not the bench): compress on the device, verify with the system zlib decoder, print
size and per-kernel device time.  Usage: python tools/dev_check.py [MiB] [levels]   (env ZB200_DEV selects
kernel variants, see zb_deflate.cu)."""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from zlib_b200 import load, binding as zb

L = load()
assert L.dll.zb200_init(0) == 0, L.last_error()
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
levels = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 6]
n = mb << 20
tag = f"DEV={os.environ.get('ZB200_DEV', '0')}"
kind = int(os.environ.get('DEV_KIND', '1'))
host = L.synth(n, kind=kind, seed=1)
src = torch.from_numpy(host).cuda()
cap = L.compress_bound(n) + 64
dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()

# edge sizes first (host path through compress2)
bad = 0
for kind in ((0, 1, 2) if not os.environ.get('DEV_NO_EDGE') else ()):
    for sz in (0, 1, 3, 100, 32769, 131072, 131073, 400001):
        d = L.synth(sz, kind=kind, seed=5).tobytes()
        for lv in levels:
            rc, z = L.compress2(d, lv)
            if rc != 0 or zlib.decompress(z) != d:
                bad += 1
                print(tag, "EDGE FAIL", kind, sz, lv, rc, flush=True)
print(tag, "edge cases bad =", bad, flush=True)

for level in levels:
    out = {}

    def run():
        out["n"] = L.deflate(src.data_ptr(), n, dst.data_ptr(), cap, level, zb.WRAP_ZLIB, s)

    run(); torch.cuda.synchronize()
    z = bytes(dst[:out["n"]].cpu().numpy())
    ok = zlib.decompress(z) == host.tobytes()
    L.profile(True)
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    rep = L.profile_report()
    L.profile(False)
    ms = e0.elapsed_time(e1) / reps
    ks = " ".join(f"{k[2:]}={v[0] / reps:.2f}" for k, v in rep.items() if v[0] / reps > 0.05)
    print(f"{tag} kind={kind} L{level}: ok={ok} {mb} MiB {ms:.2f} ms (profiled) ratio {n / out['n']:.4f} bytes {out['n']} | {ks}", flush=True)
