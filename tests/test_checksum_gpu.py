"""GPU: K5/K6 checksum kernels through the C-ABI vs the CPU oracle (bit-exact)."""
import ctypes as C
import json
import os
import random
import zlib

import numpy as np
import pytest

import zhelpers

pytestmark = pytest.mark.gpu


def test_kat(gpu_lib):
    k = json.load(open(os.path.join(zhelpers.GOLDEN, "kat.json")))
    for hx, want in k["crc32"]:
        assert gpu_lib.crc32(bytes.fromhex(hx)) == want
    for hx, want in k["adler32"]:
        assert gpu_lib.adler32(bytes.fromhex(hx)) == want
    tab = gpu_lib.dll.get_crc_table()
    assert [tab[i] for i in range(8)] == k["crc_table_0_8"]


def test_golden_stream_checksums(gpu_lib):
    for e in json.load(open(os.path.join(zhelpers.GOLDEN, "streams.json"))):
        data = zhelpers.corpus(e["kind"], e["n"], e["seed"])
        assert gpu_lib.crc32(data) == e["crc32"]
        assert gpu_lib.adler32(data) == e["adler32"]


def test_random_lengths_alignments_seeds(gpu_lib, oracle):
    rng = random.Random(11)
    big = np.frombuffer(rng.randbytes(3 << 20), dtype=np.uint8)
    edge = [0, 1, 2, 3, 15, 16, 17, 31, 32, 33, 511, 512, 513, 5551, 5552, 5553, 65535, 65536, 65537,
            (1 << 20) - 1, 1 << 20, (1 << 20) + 1, (1 << 20) + 527, (2 << 20) + 12345]
    for t in range(300):
        n = edge[t] if t < len(edge) else rng.randint(0, 2_500_000)
        off = rng.randint(0, 64)
        d = big[off:off + n]
        for seed in (0, 1, rng.getrandbits(32), 0xFFFFFFFF):
            assert gpu_lib.crc32(d, seed) == oracle.crc32(d, seed), (n, off, seed)
            assert gpu_lib.adler32(d, seed) == oracle.adler32(d, seed), (n, off, seed)


def test_worst_case_adler_bytes(gpu_lib, oracle):
    for n in (5552, 5553, 70000, 1 << 20, (5 << 20) + 7):
        d = np.full(n, 0xFF, dtype=np.uint8)
        assert gpu_lib.adler32(d) == oracle.adler32(d)
        assert gpu_lib.crc32(d) == oracle.crc32(d)


def test_device_pointers_and_combine(gpu_lib, oracle):
    import torch
    n = (48 << 20) + 4321
    host = gpu_lib.synth(n, kind=1, seed=5)
    dev = torch.from_numpy(host).cuda()
    for off, ln in ((0, n), (1, n - 1), (13, 1 << 20), (16, (32 << 20) + 5), (4097, 3)):
        crc, adl = gpu_lib.checksum(dev.data_ptr() + off, ln)
        assert crc == oracle.crc32(host[off:off + ln]) and adl == oracle.adler32(host[off:off + ln]), (off, ln)
    # zlib.h entry points accept device pointers too
    assert gpu_lib.crc32(dev.data_ptr(), 0, 1 << 20) == oracle.crc32(host[:1 << 20])
    # checksum of checksums: slices + _combine == whole (size-independent property)
    cuts = [0, 5 << 20, (17 << 20) + 3, (40 << 20) + 1, n]
    crc_all, adl_all = gpu_lib.checksum(dev.data_ptr(), n)
    crc, adl = 0, 1
    for a, b in zip(cuts, cuts[1:]):
        c, ad = gpu_lib.checksum(dev.data_ptr() + a, b - a)
        crc = gpu_lib.crc32_combine(crc, c, b - a)
        adl = gpu_lib.adler32_combine(adl, ad, b - a)
    assert crc == crc_all and adl == adl_all
    # device-resident result form
    out = torch.zeros(2, dtype=torch.int32, device="cuda")
    gpu_lib.checksum_dev(dev.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream())
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint32)
    assert int(got[0]) == crc_all and int(got[1]) == adl_all


def test_full_size_property(gpu_lib, oracle):
    """1 GiB: halves + combine == whole, and a 64 MiB prefix equals the oracle."""
    import torch
    n = 1 << 30
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    tile = torch.from_numpy(gpu_lib.synth(64 << 20, kind=1, seed=9)).cuda()
    for i in range(16):
        dev[i * (64 << 20):(i + 1) * (64 << 20)] = tile ^ i
    whole = gpu_lib.checksum(dev.data_ptr(), n)
    h = n // 2 + 77
    a = gpu_lib.checksum(dev.data_ptr(), h)
    b = gpu_lib.checksum(dev.data_ptr() + h, n - h)
    assert gpu_lib.crc32_combine(a[0], b[0], n - h) == whole[0]
    assert gpu_lib.adler32_combine(a[1], b[1], n - h) == whole[1]
    pre = tile.cpu().numpy()
    assert gpu_lib.checksum(dev.data_ptr(), 64 << 20) == (oracle.crc32(pre), oracle.adler32(pre))


def test_checksum_batch(gpu_lib, oracle):
    """zb200_checksum_batch: per-buffer crc32 / adler32 of n buffers in one launch (empty, tiny, unaligned, large)."""
    import random
    rng = random.Random(4)
    sizes = [0, 1, 2, 3, 15, 16, 17, 255, 256, 257, 4095, 4096, 4097, 65521, 65522, 100001, 1 << 20, (3 << 20) + 7] + \
            [rng.randint(0, 50000) for _ in range(40)]
    bufs = [rng.randbytes(n) for n in sizes]
    crcs, adls = gpu_lib.checksum_batch(bufs)
    for b, c, a in zip(bufs, crcs, adls):
        assert c == oracle.crc32(b) and a == oracle.adler32(b), len(b)
    bufs = [bytes([255]) * 70000, bytes(70000)]                # extreme sums
    crcs, adls = gpu_lib.checksum_batch(bufs)
    assert crcs == [oracle.crc32(b) for b in bufs] and adls == [oracle.adler32(b) for b in bufs]


def test_ten_thousand_unaligned_slices(gpu_lib, oracle):
    """SURVEY.md 8(d) gate (c): bit-equal crc32 / adler32 on >= 10^4 (offset, length) pairs, unaligned starts and ends --
    consecutive random-length slices of one buffer through zb200_checksum_batch (every start alignment mod 16 occurs),
    then the whole buffer rebuilt from the slices with the library's crc32_combine and the exact Adler join (the
    reference's adler32_combine keeps a non-canonical 65521 now and then, adler32.c:142-147, so it is not chained here)."""
    from zlib_b200.dist import adler_join
    import random
    rng = random.Random(2026)
    sizes = [rng.choice([0, 1, 2, 3, 5, 15, 16, 17, 31, 33, rng.randint(0, 300), rng.randint(0, 9000)]) for _ in range(10000)]
    sizes[1234] = 300001
    blob = zhelpers.corpus(0, sum(sizes), 3)
    bufs, pos = [], 0
    for n in sizes:
        bufs.append(blob[pos:pos + n])
        pos += n
    crcs, adls = gpu_lib.checksum_batch(bufs)
    crc_all, adl_all = 0, 1
    for b, c, a in zip(bufs, crcs, adls):
        assert c == zlib.crc32(b) and a == zlib.adler32(b), len(b)
        crc_all = gpu_lib.crc32_combine(crc_all, c, len(b))
        adl_all = adler_join(adl_all, a, len(b))
    for i in rng.sample(range(len(bufs)), 300):                    # the CPU oracle on a sample (system zlib checked all)
        assert crcs[i] == oracle.crc32(bufs[i]) and adls[i] == oracle.adler32(bufs[i])
    assert crc_all == oracle.crc32(blob) and adl_all == oracle.adler32(blob)
    assert gpu_lib.checksum(blob) == (crc_all, adl_all)
