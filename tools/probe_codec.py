"""Device-side timing of deflate / inflate (development probe, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from zlib_b200 import load, binding as zb
L = load()
assert L.dll.zb200_init(0) == 0, L.last_error()
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mb << 20
host = L.synth(n, kind=1, seed=1)
src = torch.from_numpy(host).cuda()
cap = L.compress_bound(n) + 64
dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for level in (1, 6):
    out = {}
    def run():
        out["n"] = L.deflate(src.data_ptr(), n, dst.data_ptr(), cap, level, zb.WRAP_ZLIB, s)
    ms = timed(run)
    print(f"deflate L{level}: {mb} MiB in {ms:.2f} ms = {n/ms/1e6:.2f} GB/s  ratio {n/out['n']:.3f}", flush=True)
# inflate: 64 KiB zlib streams made by the GPU deflate at level 6 (stand-in for reference streams here)
sz = 65536; ns = n // sz
import zlib
t0 = time.time()
zs = [zlib.compress(host[i*sz:(i+1)*sz].tobytes(), 6) for i in range(min(ns, 4096))]
reps = (ns + len(zs) - 1) // len(zs)
zs = (zs * reps)[:ns]
print("cpu zlib prep s", time.time() - t0)
src_off = np.zeros(ns + 1, dtype=np.int64); src_off[1:] = np.cumsum([len(z) for z in zs])
dst_off = np.arange(ns + 1, dtype=np.int64) * sz
d_src = torch.from_numpy(np.frombuffer(b"".join(zs) + b"\0" * 8, dtype=np.uint8).copy()).cuda()
d_so, d_do = torch.from_numpy(src_off).cuda(), torch.from_numpy(dst_off).cuda()
d_dst = torch.zeros(ns * sz, dtype=torch.uint8, device="cuda")
d_len = torch.zeros(ns, dtype=torch.int64, device="cuda"); d_st = torch.zeros(ns, dtype=torch.int32, device="cuda")
def runi():
    L.inflate_batch_dev(d_src.data_ptr(), d_so.data_ptr(), ns, d_dst.data_ptr(), d_do.data_ptr(), d_len.data_ptr(), d_st.data_ptr(), zb.WRAP_ZLIB, s)
ms = timed(runi)
assert int(d_st.abs().sum()) == 0
print(f"inflate: {ns} x 64 KiB in {ms:.2f} ms = {ns*sz/ms/1e6:.2f} GB/s (comp {int(src_off[-1])/ (ns*sz):.3f})")
