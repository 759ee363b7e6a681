"""ncu target: a few launches of the fused checksum over 1 GiB (development probe)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zlib_b200 import load
L = load()
assert L.dll.zb200_init(0) == 0, L.last_error()
n = 1 << 30
x = torch.empty(n, dtype=torch.uint8, device="cuda")
x.view(torch.int64).random_()
out = torch.zeros(2, dtype=torch.int32, device="cuda")
for _ in range(4):
    L.checksum_dev(x.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream())
torch.cuda.synchronize()
print(out.tolist())
