"""Development probe: zb200_deflate_batch / zb200_zip_build on many files of log-uniform size (BASELINE config 5 shape)."""
import os, sys, time, zlib, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zlib_b200 import load, binding as zb
L = load()
assert L.dll.zb200_init(0) == 0
nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
hi = float(sys.argv[2]) if len(sys.argv) > 2 else 8          # log2(max/4KiB)
rng = random.Random(1)
sizes = [int(4096 * 2 ** rng.uniform(0, hi)) for _ in range(nfiles)]
total = sum(sizes)
src = L.synth(total, kind=1, seed=3)
off = np.zeros(nfiles + 1, dtype=np.uint64); off[1:] = np.cumsum(sizes, dtype=np.uint64)
caps = [L.compress_bound(s) + 16 for s in sizes]
doff = np.zeros(nfiles + 1, dtype=np.uint64); doff[1:] = np.cumsum(caps, dtype=np.uint64)
dst = np.zeros(int(doff[-1]) + 8, dtype=np.uint8)
dlen = np.zeros(nfiles, dtype=np.uint64); crc = np.zeros(nfiles, dtype=np.uint32); st = np.zeros(nfiles, dtype=np.int32)
for rep in range(3):
    t0 = time.perf_counter()
    rc = L.dll.zb200_deflate_batch(src.ctypes.data, off.ctypes.data, nfiles, dst.ctypes.data, doff.ctypes.data, dlen.ctypes.data,
                                   crc.ctypes.data, None, st.ctypes.data, 1, zb.WRAP_RAW, None)
    dt = time.perf_counter() - t0
    assert rc == 0 and not st.any()
    print(f"deflate_batch: {nfiles} files, {total / 1e6:.0f} MB in {dt * 1e3:.1f} ms = {total / dt / 1e9:.2f} GB/s (pageable host arenas), ratio {total / dlen.sum():.3f}", flush=True)
for i in rng.sample(range(nfiles), 20):
    a, b = int(off[i]), int(off[i + 1])
    z = bytes(dst[int(doff[i]):int(doff[i]) + int(dlen[i])])
    assert zlib.decompress(z, -15) == src[a:b].tobytes() and zlib.crc32(src[a:b].tobytes()) == crc[i]
print("spot check ok")
import torch
# pinned arenas, then device arenas
for kind in ("pinned", "device"):
    if kind == "pinned":
        tsrc = torch.from_numpy(src).pin_memory(); tdst = torch.empty(len(dst), dtype=torch.uint8).pin_memory()
    else:
        tsrc = torch.from_numpy(src).cuda(); tdst = torch.empty(len(dst), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        rc = L.dll.zb200_deflate_batch(tsrc.data_ptr(), off.ctypes.data, nfiles, tdst.data_ptr(), doff.ctypes.data, dlen.ctypes.data,
                                       crc.ctypes.data, None, st.ctypes.data, 1, zb.WRAP_RAW, None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert rc == 0 and not st.any()
        print(f"deflate_batch ({kind} arenas): {dt * 1e3:.1f} ms = {total / dt / 1e9:.2f} GB/s", flush=True)
    out = tdst.cpu().numpy()
    for i in rng.sample(range(nfiles), 10):
        a, b = int(off[i]), int(off[i + 1])
        assert zlib.decompress(bytes(out[int(doff[i]):int(doff[i]) + int(dlen[i])]), -15) == src[a:b].tobytes()
L.dll.zb200_profile(1)
rc = L.dll.zb200_deflate_batch(tsrc.data_ptr(), off.ctypes.data, nfiles, tdst.data_ptr(), doff.ctypes.data, dlen.ctypes.data,
                               crc.ctypes.data, None, st.ctypes.data, 1, zb.WRAP_RAW, None)
torch.cuda.synchronize()
print(L.profile_report() if hasattr(L, "profile_report") else "")
