"""ncu target: two level-6 passes over 128 MiB of the mixed corpus (development probe)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zlib_b200 import load, binding as zb, synth
L = load(); assert L.dll.zb200_init(0) == 0
n = 128 << 20
d = torch.from_numpy(synth.synth(n, 1, 1)).cuda()
cap = L.compress_bound(n) + 64
o = torch.empty(cap, dtype=torch.uint8, device="cuda")
for _ in range(2):
    clen = L.deflate(d.data_ptr(), n, o.data_ptr(), cap, 6, zb.WRAP_ZLIB, torch.cuda.current_stream())
torch.cuda.synchronize()
print(clen)
