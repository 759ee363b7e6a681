timeout 900 python -m pytest tests/test_checksum_gpu.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, time, zlib; sys.path.insert(0, '.')
from zlib_b200 import load
L = load(); assert L.dll.zb200_init(0) == 0
x = L.synth(1 << 30, kind=2, seed=5)
for _ in range(3):
    t0 = time.perf_counter(); c = L.crc32(x); dt = time.perf_counter() - t0
    print(f"crc32() of a 1 GiB malloc'ed buffer: {dt*1e3:.1f} ms = {len(x)/dt/1e9:.1f} GB/s")
t0 = time.perf_counter(); w = zlib.crc32(x); print("system zlib one core:", round(time.perf_counter() - t0, 2), "s", c == w)
PY
