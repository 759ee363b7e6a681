"""CPU: pin the oracle (oracle/zoracle.c) to the reference.

1. against the committed golden fixtures (outputs of the reference itself, produced by
   tests/golden/make_golden.py), always;
2. differentially against oracle/_ref/libzref.so (the unmodified reference compiled by
   oracle/Makefile) whenever that build is present, including byte-identical deflate output.
"""
import base64
import json
import os
import random

import pytest

import zhelpers

G = zhelpers.GOLDEN


def _load(name):
    return json.load(open(os.path.join(G, name)))


def test_kat_checksums(oracle):
    k = _load("kat.json")
    for hx, want in k["crc32"]:
        assert oracle.crc32(bytes.fromhex(hx)) == want
    for hx, want in k["adler32"]:
        assert oracle.adler32(bytes.fromhex(hx)) == want
    assert oracle.dll.zo_crc32(0, None, 0) == k["crc32_null"] == 0
    assert oracle.dll.zo_adler32(0, None, 0) == k["adler32_null"] == 1
    for c1, c2, n, want in k["crc32_combine"]:
        assert oracle.crc32_combine(c1, c2, n) == want
    for a1, a2, n, want in k["adler32_combine"]:
        assert oracle.adler32_combine(a1, a2, n) == want
    # the SURVEY anchors themselves
    assert oracle.crc32(b"123456789") == 0xCBF43926
    assert oracle.adler32(b"123456789") == 0x091E01DE
    assert oracle.adler32(b"Wikipedia") == 0x11E60398


def test_kat_compress2_bytes(oracle):
    for hx, level, want in _load("kat.json")["compress2"]:
        data = bytes.fromhex(hx)
        assert oracle.deflate(data, level).hex() == want
        rc, out, used = oracle.inflate(bytes.fromhex(want), len(data))
        assert rc == 0 and out == data and used == len(want) // 2


def test_golden_streams(oracle):
    for e in _load("streams.json"):
        data = zhelpers.corpus(e["kind"], e["n"], e["seed"])
        assert oracle.crc32(data) == e["crc32"]
        assert oracle.adler32(data) == e["adler32"]
        for level, z in e["z"].items():
            z = base64.b64decode(z)
            assert oracle.deflate(data, int(level)) == z, (e["kind"], e["n"], level)
            rc, out, used = oracle.inflate(z, e["n"])
            assert rc == 0 and out == data and used == len(z)


def test_golden_corrupt(oracle):
    for e in _load("corrupt.json"):
        z = base64.b64decode(e["z"])
        rc, out, _ = oracle.inflate(z, e["cap"])
        assert rc == e["rc"], (e["kind"], e["n"], e["seed"], rc, e["rc"], oracle.last_msg())
        if e["out_ok"]:
            assert out == zhelpers.corpus(e["kind"], e["n"], e["seed"])


def test_gzip_and_raw_wrappers(oracle):
    import zlib
    data = zhelpers.corpus(1, 30000, 5)
    for wrap, wbits in ((0, -15), (2, 31)):
        z = oracle.deflate(data, 6, wrap)
        assert zlib.decompress(z, wbits) == data            # independent decoder (system zlib)
        rc, out, used = oracle.inflate(z, len(data), wrap)
        assert rc == 0 and out == data and used == len(z)
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    z = co.compress(data) + co.flush()
    rc, out, _ = oracle.inflate(z, len(data), 2)
    assert rc == 0 and out == data


def test_differential_checksums(oracle, ref):
    rng = random.Random(1)
    for t in range(400):
        n = rng.choice([0, 1, 2, 3, 15, 16, 17, 5551, 5552, 5553, rng.randint(0, 70000)])
        d = rng.randbytes(n + 3)[rng.randint(0, 3):][:n]
        seed = rng.choice([0, 1, 0xFFFFFFFF, rng.getrandbits(32)])
        assert ref.crc32(d, seed) == oracle.crc32(d, seed)
        assert ref.adler32(d, seed) & 0xFFFFFFFF == oracle.adler32(d, seed)
        l2 = rng.choice([0, 1, 2, 65520, 65521, 65522, rng.getrandbits(20), rng.getrandbits(40)])
        c1, c2 = rng.getrandbits(32), rng.getrandbits(32)
        assert ref.dll.crc32_combine(c1, c2, l2) == oracle.crc32_combine(c1, c2, l2)
        a1 = (rng.randrange(65521) << 16) | rng.randrange(65521)
        a2 = (rng.randrange(65521) << 16) | rng.randrange(65521)
        assert ref.dll.adler32_combine(a1, a2, l2) == oracle.adler32_combine(a1, a2, l2)


def test_differential_deflate_bytes(oracle, ref):
    rng = random.Random(2)
    for t in range(24):
        kind = rng.randrange(5)
        n = rng.choice([0, 1, 2, 3, 4, 5, 10, 100, 1000, 32767, 32768, 65535, 65536, 65537,
                        rng.randint(0, 300000), rng.randint(0, 300000)])
        d = zhelpers.corpus(kind, n, t)
        for level in (range(10) if t % 4 == 0 else (1, 6)):
            a, b = ref.compress2(d, level), oracle.deflate(d, level)
            assert a == b, (kind, n, level, len(a), len(b))
            rc, out, _ = oracle.inflate(a, n)
            assert rc == 0 and out == d


def test_differential_inflate_errors(oracle, ref):
    rng = random.Random(3)
    for t in range(1500):
        kind = rng.randrange(5)
        n = rng.choice([10, 200, 3000, 40000])
        d = zhelpers.corpus(kind, n, t)
        c = bytearray(ref.compress2(d, rng.choice([0, 1, 6, 9])))
        mode = rng.randrange(4)
        if mode == 0:
            c[rng.randrange(len(c))] ^= 1 << rng.randrange(8)
        elif mode == 1:
            c = c[:rng.randrange(len(c))]
        elif mode == 2:
            for _ in range(3):
                c[rng.randrange(len(c))] = rng.getrandbits(8)
        cap = rng.choice([n, n, n, n // 2, n + 100, 0])
        ra, oa = ref.uncompress(bytes(c), cap)
        rb, ob, _ = oracle.inflate(bytes(c), cap)
        assert ra == rb, (t, kind, n, mode, cap, ra, rb)
        if ra == 0:
            assert oa == ob
