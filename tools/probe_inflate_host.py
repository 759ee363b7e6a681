"""zb200_inflate_batch with pinned host arenas (development probe): H2D + decode + D2H per call."""
import sys, os, time, zlib, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zlib_b200 import load, binding as zb
L = load()
assert L.dll.zb200_init(0) == 0
sz, distinct = 65536, 1024
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
host = L.synth(distinct * sz, kind=1, seed=1)
zs = [zlib.compress(host[i * sz:(i + 1) * sz].tobytes(), 6) for i in range(distinct)]
zs = (zs * ((ns + distinct - 1) // distinct))[:ns]
so = np.zeros(ns + 1, dtype=np.uint64); so[1:] = np.cumsum([len(z) for z in zs], dtype=np.uint64)
do = np.arange(ns + 1, dtype=np.uint64) * sz
src = torch.from_numpy(np.frombuffer(b"".join(zs) + b"\0" * 8, dtype=np.uint8).copy()).pin_memory()
dst = torch.empty(ns * sz, dtype=torch.uint8).pin_memory()
dl = np.zeros(ns, dtype=np.uint64); st = np.zeros(ns, dtype=np.int32)
for rep in range(4):
    t0 = time.perf_counter()
    rc = L.dll.zb200_inflate_batch(src.data_ptr(), so.ctypes.data, ns, dst.data_ptr(), do.ctypes.data, dl.ctypes.data, st.ctypes.data, zb.WRAP_ZLIB, None)
    dt = time.perf_counter() - t0
    assert rc == 0 and not st.any() and (dl == sz).all()
    print(f"inflate_batch pinned host arenas: {ns} x 64 KiB in {dt * 1e3:.1f} ms = {ns * sz / dt / 1e9:.2f} GB/s (in {int(so[-1]) / 1e6:.0f} MB)", flush=True)
src_p = src.numpy().copy(); dst_p = np.zeros(ns * sz, dtype=np.uint8)            # pageable arenas
for rep in range(3):
    t0 = time.perf_counter()
    rc = L.dll.zb200_inflate_batch(src_p.ctypes.data, so.ctypes.data, ns, dst_p.ctypes.data, do.ctypes.data, dl.ctypes.data, st.ctypes.data, zb.WRAP_ZLIB, None)
    dt = time.perf_counter() - t0
    assert rc == 0 and not st.any() and (dl == sz).all()
    print(f"inflate_batch pageable host arenas: {dt * 1e3:.1f} ms = {ns * sz / dt / 1e9:.2f} GB/s", flush=True)
assert np.array_equal(dst_p, dst.numpy())
out = dst.numpy()
for i in (0, 1, distinct - 1, distinct, ns - 1):
    assert out[i * sz:(i + 1) * sz].tobytes() == host[(i % distinct) * sz:(i % distinct + 1) * sz].tobytes()
print("ok")
