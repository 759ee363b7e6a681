"""Quick device-side timing of the checksum kernels (development probe, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zlib_b200 import load
L = load()
assert L.dll.zb200_init(0) == 0, L.last_error()
for n in (1 << 20, 16 << 20, 256 << 20, 1 << 30, 4 << 30):
    x = torch.randint(0, 255, (n,), dtype=torch.uint8, device="cuda")
    out = torch.zeros(2, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    for _ in range(3):
        L.checksum_dev(x.data_ptr(), n, out.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        L.checksum_dev(x.data_ptr(), n, out.data_ptr(), s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"n={n>>20} MiB  {ms:.3f} ms  {n/ms/1e6:.1f} GB/s")
    del x
