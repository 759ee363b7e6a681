#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the compiled reference (oracle/_ref/libzref.so).

Run in the dev container (needs /root/reference to have been built by oracle/Makefile):
    python tests/golden/make_golden.py
The fixtures pin the oracle and the CUDA path to outputs of the reference itself:
  kat.json      known-answer checksums / compress2 bytes quoted in SURVEY.md 8(c)
  streams.json  reference compress2 output for seeded inputs (levels 0,1,6,9), with the
                checksums of the inputs
  corrupt.json  damaged streams with the return code of the reference's uncompress()
Inputs are described by (kind, n, seed) of tests/zhelpers.corpus, not stored.
"""
import base64
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import zhelpers  # noqa: E402


def b64(b):
    return base64.b64encode(b).decode()


def main():
    ref = zhelpers.Ref()
    kat = {
        "crc32": [["313233343536373839", ref.crc32(b"123456789")], ["", ref.crc32(b"")]],
        "adler32": [["313233343536373839", ref.adler32(b"123456789")], ["57696b697065646961", ref.adler32(b"Wikipedia")],
                    ["", ref.adler32(b"")]],
        "crc32_null": ref.dll.crc32(0, None, 0),
        "adler32_null": ref.dll.adler32(0, None, 0),
        "crc32_combine": [[ref.crc32(b"12345"), ref.crc32(b"6789"), 4,
                           ref.dll.crc32_combine(ref.crc32(b"12345"), ref.crc32(b"6789"), 4)]],
        "adler32_combine": [[ref.adler32(b"12345"), ref.adler32(b"6789"), 4,
                             ref.dll.adler32_combine(ref.adler32(b"12345"), ref.adler32(b"6789"), 4)]],
        "compress2": [],
        "version": ref.dll.zlibVersion().decode(),
        "compile_flags": ref.dll.zlibCompileFlags(),
        "crc_table_0_8": [ref.dll.get_crc_table()[i] for i in range(8)],
    }
    rng = random.Random(7)
    for _ in range(40):
        c1, c2 = rng.getrandbits(32), rng.getrandbits(32)
        n = rng.choice([1, 2, 3, 255, 65520, 65521, 65522, rng.getrandbits(20), rng.getrandbits(33)])
        kat["crc32_combine"].append([c1, c2, n, ref.dll.crc32_combine(c1, c2, n)])
        a1 = (rng.randrange(65521) << 16) | rng.randrange(65521)
        a2 = (rng.randrange(65521) << 16) | rng.randrange(65521)
        kat["adler32_combine"].append([a1, a2, n, ref.dll.adler32_combine(a1, a2, n)])
    hello = b"hello, hello!\0"
    for data in (hello, b"", bytes(100)):
        for level in (0, 1, 6, 9):
            kat["compress2"].append([data.hex(), level, ref.compress2(data, level).hex()])
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

    streams = []
    cases = [(k, n, s) for s, (k, n) in enumerate(
        [(1, 0), (1, 1), (1, 2), (1, 3), (0, 7), (1, 100), (4, 1000), (1, 5000), (3, 5000), (0, 5000), (2, 5000),
         (1, 32767), (1, 32768), (4, 40000), (3, 65536), (1, 65537), (0, 70000), (2, 100000), (1, 150000), (3, 140000)])]
    for kind, n, seed in cases:
        data = zhelpers.corpus(kind, n, seed)
        entry = {"kind": kind, "n": n, "seed": seed, "crc32": ref.crc32(data), "adler32": ref.adler32(data), "z": {}}
        for level in (0, 1, 6, 9):
            if level == 0 and n > 40000:
                continue                      # stored copies of big inputs only bloat the fixture
            entry["z"][str(level)] = b64(ref.compress2(data, level))
        streams.append(entry)
    json.dump(streams, open(os.path.join(HERE, "streams.json"), "w"))

    corrupt = []
    for i in range(160):
        kind, n = rng.choice([(1, 300), (3, 3000), (1, 20000), (0, 2000), (4, 9000)])
        seed = 100 + i
        data = zhelpers.corpus(kind, n, seed)
        level = rng.choice([0, 1, 6, 9])
        c = bytearray(ref.compress2(data, level))
        mode = i % 4
        if mode == 0:
            j = rng.randrange(len(c)); c[j] ^= 1 << rng.randrange(8)
        elif mode == 1:
            c = c[:rng.randrange(len(c))]
        elif mode == 2:
            for _ in range(3):
                c[rng.randrange(len(c))] = rng.getrandbits(8)
        cap = rng.choice([n, n, n, n // 2, n + 100, 0])
        rc, out = ref.uncompress(bytes(c), cap)
        corrupt.append({"kind": kind, "n": n, "seed": seed, "cap": cap, "z": b64(bytes(c)), "rc": rc,
                        "out_ok": bool(rc == 0 and out == data)})
    json.dump(corrupt, open(os.path.join(HERE, "corrupt.json"), "w"))
    print("kat", len(kat["compress2"]), "streams", len(streams), "corrupt", len(corrupt))


if __name__ == "__main__":
    main()
