"""CPU (gloo, world_size 2 and 3): the multi-GPU host logic -- shard ranges, the all-gather of
{len, n, crc, adler}, checksum combination, variable-length gather and stream framing.
The GPU compressor cannot run here, so a test-only stand-in (system zlib with a preset
dictionary and Z_SYNC_FLUSH) produces the per-rank shards; the assembled stream is then decoded
by the CPU oracle."""
import os
import socket
import sys
import zlib

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _standin_compress(data, halo, last, out):
    raw = bytes(data.numpy())
    zd = bytes(halo.numpy()) if halo is not None and halo.numel() else None
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, zd) if zd else zlib.compressobj(6, zlib.DEFLATED, -15)
    z = co.compress(raw) + (co.flush(zlib.Z_FINISH) if last else co.flush(zlib.Z_SYNC_FLUSH))
    out[:len(z)] = torch.frombuffer(bytearray(z), dtype=torch.uint8)
    return len(z), zlib.crc32(raw), zlib.adler32(raw)


def _worker(rank, world, port, n, wrap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import zhelpers
        from zlib_b200 import dist as zd, load, binding as zb
        lib = load()
        data = zhelpers.corpus(1, n, 5)
        ranges = zd.shard_ranges(n, world)
        a, b = ranges[rank]
        mine = torch.frombuffer(bytearray(data[a:b]), dtype=torch.uint8) if b > a else torch.empty(0, dtype=torch.uint8)
        halo = torch.frombuffer(bytearray(data[max(0, a - zd.WINDOW):a]), dtype=torch.uint8) if a > 0 else None
        plan, out, clen, full = zd.deflate_sharded(lib, mine, halo, 6, wrap, compress_fn=_standin_compress)
        assert plan.n_in == n
        assert plan.adler32 == zlib.adler32(data) and plan.crc32 == zlib.crc32(data)
        if rank == 0:
            orc = zhelpers.Oracle()
            z = bytes(full.numpy())
            assert len(z) == plan.total
            rc, dec, used = orc.inflate(z, n, wrap)
            assert rc == 0 and dec == data and used == len(z), (rc, used, len(z))
        else:
            assert full is None
        q.put((rank, "ok"))
    except Exception as e:          # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,wrap", [(2, 1_000_003, 1), (3, 131072 * 2 + 5, 2), (2, 100, 1)])
def test_sharded_stream_assembly(world, n, wrap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, wrap, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(m == "ok" for _, m in res), res


def test_shard_ranges_and_adler_join():
    from zlib_b200 import dist as zd
    for n in (0, 1, 131072, 131073, 10 * 131072 + 7):
        for w in (1, 2, 3, 8):
            r = zd.shard_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(a % 131072 == 0 for a, _ in r if a < n)
    import random
    rng = random.Random(1)
    for _ in range(200):
        x, y = rng.randbytes(rng.randint(0, 70000)), rng.randbytes(rng.randint(0, 70000))
        assert zd.adler_join(zlib.adler32(x), zlib.adler32(y), len(y)) == zlib.adler32(x + y)


def _worker_c3_c4(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import random
        import zhelpers
        from zlib_b200 import dist as zd, load
        lib = load()
        # config 4: one buffer, a slice per rank, combine in rank order (stand-in checksum: system zlib)
        data = zhelpers.corpus(0, 777_777, 9)
        a, b = zd.shard_ranges(len(data), world)[rank]
        mine = torch.frombuffer(bytearray(data[a:b]), dtype=torch.uint8) if b > a else torch.empty(0, dtype=torch.uint8)
        crc, adl, total = zd.checksum_sharded(lib, mine, checksum_fn=lambda t: (zlib.crc32(bytes(t.numpy())), zlib.adler32(bytes(t.numpy()))))
        assert (crc, adl, total) == (zlib.crc32(data), zlib.adler32(data), len(data))
        # config 3: streams sharded by compressed size; statuses and lengths known everywhere afterwards
        rng = random.Random(3)
        plain = [zhelpers.corpus(rng.randrange(5), rng.randint(0, 40000), 100 + i) for i in range(37)]
        zs = [zlib.compress(p, 6) for p in plain]
        zs[5] = zs[5][:-3]                                         # one damaged stream
        def standin(zz, caps):
            outs, st = [], []
            for z in zz:
                try:
                    outs.append(zlib.decompress(z)); st.append(0)
                except zlib.error:
                    outs.append(b""); st.append(-3)
            return outs, st
        idx, outs, st, lens = zd.inflate_sharded(lib, zs, [len(p) for p in plain], inflate_fn=standin)
        assert st == [0 if i != 5 else -3 for i in range(len(zs))]
        assert lens == [len(p) if i != 5 else 0 for i, p in enumerate(plain)]
        assert all(outs[k] == plain[i] for k, i in enumerate(idx) if i != 5)
        parts = zd.balance_streams([len(z) for z in zs], world)
        assert [i for p in parts for i in p] == list(range(len(zs)))
        q.put((rank, "ok"))
    except Exception:          # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_checksum_and_inflate(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_c3_c4, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(m == "ok" for _, m in res), res


def _standin_segment(names, datas):
    """Test-only stand-in for zb200_zip_segment: system zlib for the members, the local headers of zip.c:969-1032."""
    import struct
    from zlib_b200.binding import DOS_DATETIME
    seg, metas = bytearray(), []
    for nm, d in zip(names, datas):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        z = co.compress(d) + co.flush()
        metas.append((len(seg), len(z), len(d), zlib.crc32(d)))
        seg += struct.pack("<IHHHIIIIHH", 0x04034b50, 20, 0, 8, DOS_DATETIME, zlib.crc32(d), len(z), len(d), len(nm.encode()), 0)
        seg += nm.encode() + z
    return bytes(seg), metas


def _worker_c5(rank, world, port, q, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import io
        import random
        import zipfile
        import zhelpers
        from zlib_b200 import dist as zd, load
        lib = load()
        rng = random.Random(11)
        sizes = [0, 1, 4096, 70000, 300000] + [int(4096 * 2 ** rng.uniform(0, 7)) for _ in range(40)]
        names = [f"d{i % 3}/f{i:05d}.bin" for i in range(len(sizes))]
        datas = [zhelpers.corpus(i % 5, n, 40 + i) for i, n in enumerate(sizes)]
        shares = zd.assign_files(sizes, world)
        assert sorted(i for sh in shares for i in sh) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in sh) for sh in shares]
        assert max(loads) - min(loads) <= max(sizes)                # LPT bound
        arc = zd.zip_sharded(lib, names, datas, 6, segment_fn=_standin_segment)   # the directory is the library's own (host C)
        if rank == 0:
            zf = zipfile.ZipFile(io.BytesIO(arc))
            assert zf.testzip() is None and sorted(zf.namelist()) == sorted(names)
            for nm, d in zip(names, datas):
                assert zf.read(nm) == d
            with open(os.path.join(outdir, "sharded.zip"), "wb") as f:
                f.write(arc)
        else:
            assert arc is None
        q.put((rank, "ok"))
    except Exception:          # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_zip_archive(world, tmp_path):
    """BASELINE config 5 host logic: files dealt to the ranks, segments gathered, central directory by the library;
    the archive is read back by Python's zipfile and, when built, by the reference's miniunz."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_c5, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(m == "ok" for _, m in res), res
    miniunz = os.path.join(ROOT, "oracle", "_ref", "miniunz")
    if os.path.exists(miniunz):
        import subprocess
        out = tmp_path / "x"
        out.mkdir()
        r = subprocess.run([miniunz, "-o", str(tmp_path / "sharded.zip")], cwd=out, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert sum(len(fs) for _, _, fs in os.walk(out)) == 45


# ---------------------------------------------------------------------------------------------
# deflate_rounds: pieces dealt round robin, transfers of round j while round j + 1 is compressed (bench.py --gpus N)
# ---------------------------------------------------------------------------------------------
class _StandInLib:
    """What deflate_rounds needs from the library, on host pointers: system zlib as the per-piece compressor (raw
    deflate, preset dictionary, Z_SYNC_FLUSH unless last); the combine is the real host-side crc32_combine."""

    def __init__(self, real):
        import ctypes as C
        self.real, self.C = real, C
        outer = self

        class _Dll:
            @staticmethod
            def zb200_copy_async(dst, src, n, stream):
                C.memmove(dst, src, n)
                return 0

            @staticmethod
            def zb200_sync(stream):
                return 0
        self.dll = _Dll()

    def _check(self, rc, what):
        assert rc == 0, what

    def crc32_combine(self, a, b, n):
        return self.real.crc32_combine(a, b, n)

    def deflate_shard(self, src, n, dict_, dict_len, dst, cap, level, wrap, flags, stream=None):
        C = self.C
        raw = C.string_at(src, n)
        zd = C.string_at(dict_, dict_len) if dict_len else None
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, zd) if zd else zlib.compressobj(6, zlib.DEFLATED, -15)
        z = co.compress(raw) + (co.flush(zlib.Z_SYNC_FLUSH) if (flags & 1) else co.flush(zlib.Z_FINISH))
        assert len(z) <= cap
        C.memmove(dst, z, len(z))
        return len(z), zlib.crc32(raw), zlib.adler32(raw)


def _worker_rounds(rank, world, port, total, P, host_mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        import numpy as np
        import zhelpers
        from zlib_b200 import dist as zd, load, binding as zb
        lib = _StandInLib(load())
        data = np.frombuffer(zhelpers.corpus(1, total, 8), dtype=np.uint8).copy()
        ranges = zd.piece_ranges(total, world, P)
        assert sorted(x for r in ranges for x in r if x[1] > x[0])[0][0] == 0
        pieces, outs = [], []
        for a, b in ranges[rank]:
            dl = min(a, zd.WINDOW)
            pieces.append((data.ctypes.data + a, b - a, data.ctypes.data + a - dl, dl))
            outs.append(torch.empty(b - a + 4096, dtype=torch.uint8))
        # the last piece of the stream is the last NON-EMPTY one; here sizes are chosen so that every piece is non-empty
        final = torch.zeros(total + 65536, dtype=torch.uint8) if rank == 0 and not host_mode else None
        host_final = None
        if host_mode:                                            # every rank writes into its own buffer; rank 0's is checked piecewise
            hbuf = np.zeros(total + 65536, dtype=np.uint8)
            host_final = hbuf.ctypes.data
        tot, crc, adl, n_in = zd.deflate_rounds(lib, pieces, 6, zb.WRAP_ZLIB, root=0, outs=outs, final=final, host_final=host_final)
        assert n_in == total and crc == zlib.crc32(data.tobytes()) and adl == zlib.adler32(data.tobytes())
        if host_mode:
            # gather the per-rank host buffers: in the real run they are ONE shared mapping; here add them up
            t = torch.from_numpy(hbuf.astype(np.int32))
            dist.all_reduce(t)
            z = bytes(t.to(torch.uint8).numpy()[:tot])
            if rank == 0:
                hdr = C.string_at(host_final, 2)
                assert z[:2] == hdr
        else:
            z = bytes(final.numpy()[:tot]) if rank == 0 else None
        if rank == 0:
            assert zlib.decompress(z) == data.tobytes()
            orc = zhelpers.Oracle()
            rc, dec, used = orc.inflate(z, total, 1)
            assert rc == 0 and dec == data.tobytes() and used == len(z)
        q.put((rank, "ok"))
    except Exception:          # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total,P,host_mode", [(2, 131072 * 12, 3, False), (3, 131072 * 12 - 77, 2, False), (2, 131072 * 8, 4, True)])
def test_rounds_assemble_one_stream(world, total, P, host_mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_rounds, args=(r, world, port, total, P, host_mode, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(m == "ok" for _, m in res), res
