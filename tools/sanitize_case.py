"""Small pass over every kernel family for compute-sanitizer (memcheck): sizes kept small, results still checked."""
import sys, os, zlib, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from zlib_b200 import load, binding as zb
L = load()
assert L.dll.zb200_init(0) == 0
rng = random.Random(1)
data = L.synth((3 << 20) + 777, kind=1, seed=2).tobytes()
assert L.checksum(data) == (zlib.crc32(data), zlib.adler32(data))
for level in (0, 1, 6):
    rc, z = L.compress2(data, level)
    assert rc == 0 and zlib.decompress(z) == data
    rc, out = L.uncompress(z, len(data))                      # marker path
    assert rc == 0 and out == data
zf = zlib.compress(data, 6)                                    # block finder path
rc, out = L.uncompress(zf, len(data))
assert rc == 0 and out == data
bad = bytearray(zf); bad[len(bad) // 2] ^= 8
rc, out = L.uncompress(bytes(bad), len(data))
assert rc != 0
bufs = [data[i * 40000:(i * 40000) + rng.randint(0, 300000)] for i in range(40)] + [b"", b"x"]
for level, wrap, wb in ((1, zb.WRAP_RAW, -15), (6, zb.WRAP_GZIP, 31)):
    outs, st, crcs, adls = L.deflate_batch(bufs, level, wrap)
    assert st == [0] * len(bufs) and all(zlib.decompress(z, wb) == b for z, b in zip(outs, bufs))
    assert crcs == [zlib.crc32(b) for b in bufs]
zs = [zlib.compress(b, 6) for b in bufs]
outs, st = L.inflate_batch(zs, [len(b) for b in bufs])
assert st == [0] * len(bufs) and outs == bufs
arc = L.zip_build({f"f{i}.bin": b for i, b in enumerate(bufs)}, level=1)
import io, zipfile
assert zipfile.ZipFile(io.BytesIO(arc)).testzip() is None
rc, out, msg, tin = L.inflate_stream(zf, 15, 70000, 50000, zb.Z_NO_FLUSH)
assert rc == zb.Z_STREAM_END and out == data
rc, zz = L.deflate_stream(data[:500000], level=6, wbits=15)
assert rc == 0 and zlib.decompress(zz) == data[:500000]
print("sanitize case ok")
