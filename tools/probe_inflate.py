"""ncu target: batch inflate of 8192 zlib streams of 64 KiB (level 6, system zlib) -- development probe."""
import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zlib_b200 import load, binding as zb, synth
L = load()
assert L.dll.zb200_init(0) == 0, L.last_error()
sz, distinct, ns = 65536, 512, 8192
base = synth.synth(distinct * sz, 1, 1)
zs = [zlib.compress(base[i * sz:(i + 1) * sz].tobytes(), 6) for i in range(distinct)]
zs = (zs * (ns // distinct))[:ns]
off = np.zeros(ns + 1, dtype=np.int64); off[1:] = np.cumsum([len(z) for z in zs])
d_z = torch.from_numpy(np.frombuffer(b"".join(zs) + b"\0" * 8, dtype=np.uint8).copy()).cuda()
d_so = torch.from_numpy(off).cuda()
d_do = torch.arange(ns + 1, dtype=torch.int64, device="cuda") * sz
d_out = torch.empty(ns * sz, dtype=torch.uint8, device="cuda")
d_len = torch.zeros(ns, dtype=torch.int64, device="cuda"); d_st = torch.zeros(ns, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream()
for _ in range(3):
    L.inflate_batch_dev(d_z.data_ptr(), d_so.data_ptr(), ns, d_out.data_ptr(), d_do.data_ptr(), d_len.data_ptr(), d_st.data_ptr(), zb.WRAP_ZLIB, s)
torch.cuda.synchronize()
assert int(d_st.abs().sum()) == 0
print("ok")
