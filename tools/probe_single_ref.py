"""uncompress() of one long stream WITHOUT flush points (system zlib = the reference's format), config 1 shape: 64 MiB of text, level 6."""
import sys, os, time, zlib, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zlib_b200 import load
L = load()
assert L.dll.zb200_init(0) == 0
for mb, kind in ((64, 0), (256, 1)):
    n = mb << 20
    data = L.synth(n, kind=kind, seed=1).tobytes()
    t0 = time.perf_counter(); z = zlib.compress(data, 6); tc = time.perf_counter() - t0
    t0 = time.perf_counter(); zlib.decompress(z); td = time.perf_counter() - t0
    print(f"{mb} MiB kind {kind}: system zlib level 6 {tc:.2f} s -> {len(z) >> 20} MiB; its own decompress {td * 1e3:.0f} ms ({n / td / 1e9:.2f} GB/s, one core)")
    out = C.create_string_buffer(n)
    for rep in range(3):
        ul = C.c_ulong(n)
        t0 = time.perf_counter(); rc = L.dll.uncompress(out, C.byref(ul), z, len(z)); dt = time.perf_counter() - t0
        assert rc == 0 and ul.value == n
        print(f"   zb200 uncompress (pageable buffers): {dt * 1e3:.1f} ms = {n / dt / 1e9:.2f} GB/s")
    assert out.raw == data
    L.profile(True); ul = C.c_ulong(n); L.dll.uncompress(out, C.byref(ul), z, len(z)); print("  ", {k: round(v[0], 2) for k, v in L.profile_report().items()}); L.profile(False)
