"""compress2 / uncompress with pageable host buffers (what an unmodified caller of the reference has), vs registering them."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zlib_b200 import load
L = load()
assert L.dll.zb200_init(0) == 0
n = 1 << 30
src = L.synth(n, kind=1, seed=1)            # numpy, pageable
cap = L.compress_bound(n) + 64
dst = np.empty(cap, dtype=np.uint8)
dst[:] = 0
def comp():
    ol = C.c_ulong(cap)
    rc = L.dll.compress2(C.c_void_p(dst.ctypes.data), C.byref(ol), C.c_void_p(src.ctypes.data), n, 1)
    assert rc == 0
    return ol.value
for _ in range(3):
    t0 = time.perf_counter(); z = comp(); dt = time.perf_counter() - t0
    print(f"compress2 pageable: {dt*1e3:.1f} ms = {n/dt/1e9:.2f} GB/s")
out = np.empty(n, dtype=np.uint8); out[:] = 0
for _ in range(2):
    ol = C.c_ulong(n)
    t0 = time.perf_counter(); rc = L.dll.uncompress(C.c_void_p(out.ctypes.data), C.byref(ol), C.c_void_p(dst.ctypes.data), z); dt = time.perf_counter() - t0
    assert rc == 0 and ol.value == n
    print(f"uncompress pageable (one stream, one warp): {dt*1e3:.1f} ms = {n/dt/1e9:.3f} GB/s")
rt = torch.cuda.cudart()
t0 = time.perf_counter(); r1 = rt.cudaHostRegister(src.ctypes.data, n, 0); r2 = rt.cudaHostRegister(dst.ctypes.data, cap, 0); dt = time.perf_counter() - t0
print(f"cudaHostRegister of {n+cap>>20} MiB: {dt*1e3:.1f} ms ({r1}, {r2})")
for _ in range(3):
    t0 = time.perf_counter(); comp(); dt = time.perf_counter() - t0
    print(f"compress2 registered: {dt*1e3:.1f} ms = {n/dt/1e9:.2f} GB/s")
t0 = time.perf_counter(); rt.cudaHostUnregister(src.ctypes.data); rt.cudaHostUnregister(dst.ctypes.data); print(f"unregister {1e3*(time.perf_counter()-t0):.1f} ms")
