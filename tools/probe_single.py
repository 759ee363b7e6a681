"""uncompress() of one long stream (development probe): this library's chunked output, device- and host-resident."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zlib_b200 import load, binding as zb
L = load()
assert L.dll.zb200_init(0) == 0
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mb << 20
src = L.synth(n, kind=1, seed=1)
cap = L.compress_bound(n) + 64
for level in (1, 6):
    pz = L.dll.zb200_alloc_pinned(cap); po = L.dll.zb200_alloc_pinned(n); pi = L.dll.zb200_alloc_pinned(n)
    C.memmove(C.c_void_p(pi), C.c_void_p(src.ctypes.data), n)
    ol = C.c_ulong(cap)
    assert L.dll.compress2(C.c_void_p(pz), C.byref(ol), C.c_void_p(pi), n, level) == 0
    z = ol.value
    for rep in range(3):
        ul = C.c_ulong(n)
        t0 = time.perf_counter(); rc = L.dll.uncompress(C.c_void_p(po), C.byref(ul), C.c_void_p(pz), z); dt = time.perf_counter() - t0
        assert rc == 0 and ul.value == n
        print(f"L{level}: uncompress of one {mb} MiB stream (pinned host buffers): {dt*1e3:.1f} ms = {n/dt/1e9:.2f} GB/s", flush=True)
    assert np.array_equal(np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint8)), shape=(n,)), src)
    L.profile(True)
    ul = C.c_ulong(n); L.dll.uncompress(C.c_void_p(po), C.byref(ul), C.c_void_p(pz), z)
    print({k: round(v[0], 2) for k, v in L.profile_report().items()})
    L.profile(False)
    for p in (pz, po, pi): L.dll.zb200_free_pinned(C.c_void_p(p))
