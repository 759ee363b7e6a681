"""GPU: K4 batch inflate through the C-ABI vs golden reference streams and the CPU oracle."""
import base64
import json
import os
import random
import zlib

import numpy as np
import pytest

import zhelpers
from zlib_b200 import binding as zb

pytestmark = pytest.mark.gpu


def _golden(name):
    return json.load(open(os.path.join(zhelpers.GOLDEN, name)))


def test_kat_streams(gpu_lib):
    """compress2 outputs of the reference quoted in SURVEY.md 8(c) decode to their inputs."""
    k = _golden("kat.json")["compress2"]
    streams = [bytes.fromhex(z) for _, _, z in k]
    datas = [bytes.fromhex(d) for d, _, _ in k]
    outs, st = gpu_lib.inflate_batch(streams, [len(d) for d in datas])
    assert st == [0] * len(k)
    assert outs == datas


def test_golden_reference_streams(gpu_lib):
    streams, datas = [], []
    for e in _golden("streams.json"):
        d = zhelpers.corpus(e["kind"], e["n"], e["seed"])
        for level, z in e["z"].items():
            streams.append(base64.b64decode(z)); datas.append(d)
    outs, st = gpu_lib.inflate_batch(streams, [len(d) for d in datas])
    assert st == [0] * len(streams)
    for o, d in zip(outs, datas):
        assert o == d
    # generous capacity gives the same result
    outs, st = gpu_lib.inflate_batch(streams, [len(d) + 1000 for d in datas])
    assert st == [0] * len(streams) and outs == datas


def test_golden_corrupt_status(gpu_lib):
    """Same return code as the reference's uncompress() on damaged / truncated / short-output cases."""
    es = _golden("corrupt.json")
    streams = [base64.b64decode(e["z"]) for e in es]
    outs, st = gpu_lib.inflate_batch(streams, [e["cap"] for e in es])
    bad = [(i, es[i]["rc"], st[i]) for i in range(len(es)) if st[i] != es[i]["rc"]]
    assert not bad, bad[:10]
    for e, o in zip(es, outs):
        if e["out_ok"]:
            assert o == zhelpers.corpus(e["kind"], e["n"], e["seed"])


def test_differential_vs_oracle(gpu_lib, oracle):
    rng = random.Random(5)
    streams, caps, want = [], [], []
    for t in range(400):
        kind = rng.randrange(5)
        n = rng.choice([0, 1, 2, 3, 10, 257, 258, 259, 4000, 32768, 32769, 70000, rng.randint(0, 200000)])
        d = zhelpers.corpus(kind, n, 1000 + t)
        z = bytearray(oracle.deflate(d, rng.choice([0, 1, 1, 6, 6, 9])))
        mode = rng.randrange(6)
        if mode == 0 and len(z) > 0:
            z[rng.randrange(len(z))] ^= 1 << rng.randrange(8)
        elif mode == 1:
            z = z[:rng.randrange(len(z) + 1)]
        cap = rng.choice([n, n, n, n // 2, n + 17, 0, max(n - 1, 0)])
        rc, out, _ = oracle.inflate(bytes(z), cap)
        streams.append(bytes(z)); caps.append(cap); want.append((rc, out))
    outs, st = gpu_lib.inflate_batch(streams, caps)
    for i, (rc, out) in enumerate(want):
        assert st[i] == rc, (i, st[i], rc, len(streams[i]), caps[i])
        if rc == 0:
            assert outs[i] == out, i


def test_raw_wrap_and_unaligned(gpu_lib, oracle):
    rng = random.Random(6)
    streams, datas = [], []
    for t in range(64):
        d = zhelpers.corpus(rng.randrange(5), rng.randint(0, 50000), 2000 + t)
        streams.append(oracle.deflate(d, 6, 0)); datas.append(d)       # odd lengths -> every alignment
    outs, st = gpu_lib.inflate_batch(streams, [len(d) for d in datas], wrap=zb.WRAP_RAW)
    assert st == [0] * 64 and outs == datas


def test_hand_made_blocks(gpu_lib, oracle):
    """Fixed-code block, stored block of length 0, multi-block streams, invalid block type."""
    import zlib
    cases = []
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    d = b"fixed huffman block " * 50
    cases.append((co.compress(d) + co.flush(), d))
    co = zlib.compressobj(6)
    parts = [b"abc" * 100, b"", b"xyz" * 1000]
    z = b""
    for p in parts:
        z += co.compress(p) + co.flush(zlib.Z_FULL_FLUSH)          # empty stored blocks in the middle
    z += co.flush()
    cases.append((z, b"".join(parts)))
    outs, st = gpu_lib.inflate_batch([c[0] for c in cases], [len(c[1]) for c in cases])
    assert st == [0, 0] and outs == [c[1] for c in cases]
    bad = b"\x78\x9c\x07"                                             # block type 3
    outs, st = gpu_lib.inflate_batch([bad], [10])
    assert st == [oracle.inflate(bad, 10)[0]] == [-3]


def test_many_streams_device_resident(gpu_lib, oracle):
    """2000 x 64 KiB slices of the mixed corpus, reference/oracle-compressed, decoded on device memory."""
    import torch
    n, sz = 2000, 65536
    host = gpu_lib.synth(n * sz, kind=1, seed=21)
    zs = [oracle.deflate(host[i * sz:(i + 1) * sz], 6) for i in range(n)]
    src_off = np.zeros(n + 1, dtype=np.int64); src_off[1:] = np.cumsum([len(z) for z in zs])
    dst_off = np.arange(n + 1, dtype=np.int64) * sz
    d_src = torch.from_numpy(np.frombuffer(b"".join(zs) + b"\0" * 8, dtype=np.uint8).copy()).cuda()
    d_so, d_do = torch.from_numpy(src_off).cuda(), torch.from_numpy(dst_off).cuda()
    d_dst = torch.zeros(n * sz, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_st = torch.full((n,), -99, dtype=torch.int32, device="cuda")
    gpu_lib.inflate_batch_dev(d_src.data_ptr(), d_so.data_ptr(), n, d_dst.data_ptr(), d_do.data_ptr(),
                              d_len.data_ptr(), d_st.data_ptr(), zb.WRAP_ZLIB, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert int(d_st.abs().sum()) == 0
    assert bool((d_len == sz).all())
    assert np.array_equal(d_dst.cpu().numpy(), host)


def _gzip_member(data, level=6, name=None, extra=None, comment=None, hcrc=False):
    """A gzip member with any mix of optional header fields (RFC 1952), body from the system zlib."""
    import struct
    import zlib
    flg = (4 if extra is not None else 0) | (8 if name is not None else 0) | (16 if comment is not None else 0) | (2 if hcrc else 0)
    h = bytes([0x1f, 0x8b, 8, flg, 1, 2, 3, 4, 0, 3])
    if extra is not None:
        h += struct.pack("<H", len(extra)) + extra
    if name is not None:
        h += name + b"\0"
    if comment is not None:
        h += comment + b"\0"
    if hcrc:
        h += struct.pack("<H", zlib.crc32(h) & 0xffff)
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    return h + body + struct.pack("<II", zlib.crc32(data), len(data) & 0xffffffff)


def test_gzip_members_batch(gpu_lib, oracle):
    """windowBits+16 / +32 decoding (inflate.c:596-759, 1099-1112): header fields skipped, FHCRC, CRC-32 and ISIZE
    verified; auto-detect takes zlib and gzip streams in one batch."""
    import zlib
    rng = random.Random(8)
    datas = [zhelpers.corpus(rng.randrange(5), n, 700 + i) for i, n in enumerate([0, 1, 5, 300, 4096, 70000, 200000, 33333])]
    members = [
        _gzip_member(datas[0]), _gzip_member(datas[1], name=b"a.txt"), _gzip_member(datas[2], extra=b"\x01\x02\x03"),
        _gzip_member(datas[3], comment=b"hello", hcrc=True), _gzip_member(datas[4], name=b"n", extra=b"", comment=b"", hcrc=True),
        _gzip_member(datas[5], 1), _gzip_member(datas[6], 9, name=b"x" * 300), _gzip_member(datas[7]),
    ]
    for m, d in zip(members, datas):
        assert zlib.decompress(m, 31) == d                       # the fixtures are valid for the system decoder
    outs, st = gpu_lib.inflate_batch(members, [len(d) for d in datas], wrap=zb.WRAP_GZIP)
    assert st == [0] * len(members) and outs == datas
    # auto-detect: zlib streams and gzip members side by side
    mixed = members[:4] + [oracle.deflate(d, 6) for d in datas[4:]]
    outs, st = gpu_lib.inflate_batch(mixed, [len(d) for d in datas], wrap=3)
    assert st == [0] * len(mixed) and outs == datas
    # damage: CRC, ISIZE, header CRC, reserved flag bit, method, magic, truncation
    def flip(b, i, v=0x01):
        b = bytearray(b); b[i] ^= v; return bytes(b)
    bad = [flip(members[5], len(members[5]) - 6), flip(members[5], len(members[5]) - 2), flip(members[3], 11),
           flip(members[5], 3, 0x20), flip(members[5], 2, 0x01), flip(members[5], 1), members[5][:-3], members[5][:7]]
    want = []
    for z in bad:
        try:
            zlib.decompress(z, 31); want.append(0)
        except zlib.error:
            want.append(-3)
    outs, st = gpu_lib.inflate_batch(bad, [len(datas[5])] * 6 + [len(datas[5])] * 2, wrap=zb.WRAP_GZIP)
    assert st == want == [-3] * len(bad), (st, want)


def _used_parallel(lib, fn):
    """Runs fn with per-kernel timing on and tells whether the segment-parallel decoder's kernels were launched."""
    lib.profile(True)
    try:
        r = fn()
        rep = lib.profile_report()
    finally:
        lib.profile(False)
    return r, any("k_inflate_segments" in k for k in rep), any("k_inflate_batch" in k for k in rep)


def test_long_single_stream_decodes_in_parallel(gpu_lib, oracle):
    """uncompress() of ONE long stream: streams with byte-aligned block boundaries (this library's chunked output; any
    stream with Z_SYNC_FLUSH / Z_FULL_FLUSH points) are cut at the 00 00 FF FF markers, the segments are validated by a
    counting pass and decoded in parallel, and matches that reach back over a segment start are resolved afterwards.
    Output is bit-exact; everything the scheme cannot take goes to the serial decoder with the same result as before."""
    n = (24 << 20) + 12345
    data = gpu_lib.synth(n, kind=1, seed=21).tobytes()
    for level in (1, 6):
        rc, z = gpu_lib.compress2(data, level)
        assert rc == zb.Z_OK
        (rc, out), par, ser = _used_parallel(gpu_lib, lambda: gpu_lib.uncompress(z, n))
        assert rc == zb.Z_OK and out == data and par and not ser
        assert zlib.decompress(z) == data
    # exact capacity fits; one byte less is Z_BUF_ERROR (the serial decoder reports it the reference's way)
    rc, out = gpu_lib.uncompress(z, n - 1)
    assert rc == zb.Z_BUF_ERROR
    # a foreign stream with flush points of both kinds, matches reaching back across them, and a stored stretch
    rng = random.Random(8)
    parts = [zhelpers.corpus(rng.choice([1, 1, 3, 4]), rng.randint(50000, 400000), 50 + i) for i in range(30)]
    parts[11] = rng.randbytes(200000)                           # incompressible: stored blocks
    co = zlib.compressobj(6)
    zf = b""
    for i, p in enumerate(parts):
        zf += co.compress(p) + co.flush(zlib.Z_FULL_FLUSH if i % 5 == 4 else zlib.Z_SYNC_FLUSH)
    zf += co.flush()
    plain = b"".join(parts)
    (rc, out), par, ser = _used_parallel(gpu_lib, lambda: gpu_lib.uncompress(zf, len(plain)))
    assert rc == zb.Z_OK and out == plain and par and not ser
    # raw deflate, same boundaries
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    zr = b"".join(co.compress(p) + co.flush(zlib.Z_SYNC_FLUSH) for p in parts) + co.flush()
    outs, st = gpu_lib.inflate_batch([zr], [len(plain)], wrap=zb.WRAP_RAW)
    assert st == [0] and outs[0] == plain


def test_long_single_stream_fallbacks(gpu_lib, oracle):
    """What the parallel scheme must leave to the serial decoder: no boundaries at all, the marker pattern occurring as
    DATA (inside stored blocks), and damage -- results and return codes are those of the reference's uncompress()."""
    data = zhelpers.corpus(0, 300000, 9)                        # noise: stored blocks only, nothing to find
    z = zlib.compress(data, 6)
    (rc, out), par, ser = _used_parallel(gpu_lib, lambda: gpu_lib.uncompress(z, len(data)))
    assert rc == zb.Z_OK and out == data and ser
    # level 0: the stream shows its input, so 00 00 FF FF in the input is a false candidate every time
    rng = random.Random(10)
    raw = bytearray(rng.randbytes(900000))
    for k in range(20000, len(raw) - 8, 40000):
        raw[k:k + 4] = b"\x00\x00\xff\xff"
    raw = bytes(raw)
    z0 = zlib.compress(raw, 0)
    assert z0.count(b"\x00\x00\xff\xff") >= 20
    rc, out = gpu_lib.uncompress(z0, len(raw))
    assert rc == zb.Z_OK and out == raw
    # the same false candidates in front of real ones
    co = zlib.compressobj(0)
    zm = co.compress(raw) + co.flush(zlib.Z_SYNC_FLUSH)
    co2 = zlib.compressobj(6, zlib.DEFLATED, -15)
    # damage in the middle of a chunked stream: the reference's code and ours agree
    n = 3 << 20
    d2 = gpu_lib.synth(n, kind=1, seed=22).tobytes()
    rc, z2 = gpu_lib.compress2(d2, 1)
    bad = bytearray(z2)
    bad[len(bad) // 2] ^= 0x55
    rc, out = gpu_lib.uncompress(bytes(bad), n)
    want, _, _ = oracle.inflate(bytes(bad), n)
    assert rc == want and rc != zb.Z_OK
    bad = bytearray(z2)
    bad[-2] ^= 1                                                # Adler-32 trailer
    rc, out = gpu_lib.uncompress(bytes(bad), n)
    assert rc == zb.Z_DATA_ERROR
    rc, out = gpu_lib.uncompress(z2[:len(z2) * 2 // 3], n)      # truncated
    assert rc == zb.Z_DATA_ERROR


def test_batch_with_a_few_long_streams(gpu_lib):
    """A call with a handful of streams: the long chunked ones take the segment-parallel decoder one by one, the rest
    (short, or without boundaries) share the batch kernel; positions, lengths and statuses line up, a damaged long stream
    fails alone."""
    datas = [gpu_lib.synth((4 << 20) + 1000 * i, kind=1, seed=60 + i).tobytes() for i in range(3)]
    datas += [zhelpers.corpus(1, 5000, 3), zhelpers.corpus(3, 700000, 4), b""]
    streams = [gpu_lib.compress2(d, 1 if i % 2 else 6)[1] for i, d in enumerate(datas[:3])]
    streams += [zlib.compress(datas[3], 6), zlib.compress(datas[4], 6), zlib.compress(b"")]
    order = [3, 0, 4, 1, 5, 2]
    zs, ds = [streams[i] for i in order], [datas[i] for i in order]
    (outs, st), par, ser = _used_parallel(gpu_lib, lambda: gpu_lib.inflate_batch(zs, [len(d) for d in ds]))
    assert st == [0] * len(zs) and outs == ds and par and ser
    bad = bytearray(zs[3])
    bad[len(bad) // 3] ^= 0x10
    outs, st = gpu_lib.inflate_batch(zs[:3] + [bytes(bad)] + zs[4:], [len(d) for d in ds])
    assert st[3] == zb.Z_DATA_ERROR and [x for i, x in enumerate(st) if i != 3] == [0] * 5
    assert [o for i, o in enumerate(outs) if i != 3] == [d for i, d in enumerate(ds) if i != 3]


def test_long_stream_without_flush_points_uses_the_block_finder(gpu_lib, oracle):
    """The reference's own one-shot output (config 1's direction (b)): no markers, so block starts are FOUND -- every bit
    position that could open a non-final dynamic block is probed, survivors are checked with the decoder's own header
    code, and the counting pass keeps the ones the true decode arrives at.  Levels 1, 6, 9 of system zlib and the CPU
    oracle's streams decode bit-exact through the parallel path."""
    n = 12 << 20
    data = gpu_lib.synth(n, kind=1, seed=41).tobytes()
    for z in (zlib.compress(data, 1), zlib.compress(data, 6), zlib.compress(data[:4 << 20], 9), oracle.deflate(data[: 6 << 20], 6)):
        want = zlib.decompress(z)
        (rc, out), par, ser = _used_parallel(gpu_lib, lambda: gpu_lib.uncompress(z, len(want)))
        assert rc == zb.Z_OK and out == want and par and not ser
    # text only (long dynamic blocks), raw wrap
    text = gpu_lib.synth(8 << 20, kind=0, seed=42).tobytes()
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    zr = co.compress(text) + co.flush()
    outs, st = gpu_lib.inflate_batch([zr], [len(text)], wrap=zb.WRAP_RAW)
    assert st == [0] and outs[0] == text
    # damage in such a stream still comes out as the reference's error
    bad = bytearray(zlib.compress(data[:3 << 20], 6))
    bad[len(bad) // 2] ^= 0x04
    rc, out = gpu_lib.uncompress(bytes(bad), 3 << 20)
    want_rc, _, _ = oracle.inflate(bytes(bad), 3 << 20)
    assert rc == want_rc and rc != zb.Z_OK


def test_long_gzip_member_decodes_in_parallel(gpu_lib):
    """A long gzip member (windowBits + 16, and + 32 = auto-detect): header walked on the host, blocks found, CRC-32 and
    ISIZE of the output checked against the trailer (inflate.c:1099-1112); a wrong CRC is the serial decoder's error."""
    import gzip
    import io
    data = gpu_lib.synth(10 << 20, kind=1, seed=51).tobytes()
    buf = io.BytesIO()
    with gzip.GzipFile(filename="corpus.bin", mode="wb", fileobj=buf, compresslevel=6, mtime=1) as f:
        f.write(data)
    member = buf.getvalue()
    for wrap in (zb.WRAP_GZIP, 3):
        (outs, st), par, ser = _used_parallel(gpu_lib, lambda: gpu_lib.inflate_batch([member], [len(data)], wrap=wrap))
        assert st == [0] and outs[0] == data and par and not ser
    bad = bytearray(member)
    bad[-6] ^= 0x80                                             # CRC-32 in the trailer
    outs, st = gpu_lib.inflate_batch([bytes(bad)], [len(data)], wrap=zb.WRAP_GZIP)
    assert st == [zb.Z_DATA_ERROR]
