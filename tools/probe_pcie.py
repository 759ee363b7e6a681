"""PCIe ceiling of the host-buffer path (development probe): H2D of 1 GiB, D2H of 0.5 GiB, both at once, and compress2."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zlib_b200 import load
L = load()
assert L.dll.zb200_init(0) == 0
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D 1 GiB {a*1e3:.2f} ms = {n/a/1e9:.1f} GB/s; D2H 0.5 GiB {b*1e3:.2f} ms = {n/2/b/1e9:.1f} GB/s; both {c*1e3:.2f} ms")
import numpy as np
src = L.synth(n, kind=1, seed=1)
h_in.numpy()[:] = src
cap = L.compress_bound(n) + 64
h_z = torch.empty(cap, dtype=torch.uint8).pin_memory()
def comp():
    ol = C.c_ulong(cap)
    rc = L.dll.compress2(C.c_void_p(h_z.data_ptr()), C.byref(ol), C.c_void_p(h_in.data_ptr()), n, 1)
    assert rc == 0
e = t(comp)
print(f"compress2 pinned: {e*1e3:.2f} ms = {n/e/1e9:.1f} GB/s")
