"""Streaming inflate() rate (development probe): one z_stream fed 1 MiB at a time, no flush points in the stream."""
import sys, os, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zlib_b200 import load, binding as zb
L = load()
assert L.dll.zb200_init(0) == 0
data = L.synth(16 << 20, kind=1, seed=1).tobytes()
z = zlib.compress(data, 6)
for chunk in (1 << 20, 16384):
    t0 = time.perf_counter()
    rc, out, msg, tin = L.inflate_stream(z, 15, chunk, chunk, zb.Z_NO_FLUSH)
    dt = time.perf_counter() - t0
    assert rc == zb.Z_STREAM_END and out == data
    print(f"inflate() in {chunk >> 10} KiB steps: {len(data) / dt / 1e6:.1f} MB/s")
