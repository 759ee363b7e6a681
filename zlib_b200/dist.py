"""Multi-GPU assembly of one DEFLATE stream (one process per GPU, torch.distributed).

The path shards by input chunk: rank r compresses a contiguous,
chunk-aligned range of the input with the 32 KiB in front of it as dictionary and ends
on a byte boundary, so the shards concatenate.  The only exchange is
  1. an all-gather of four integers per rank {compressed bytes, input bytes, crc32, adler32};
  2. (optional) the variable-length gather of the compressed shards to the root rank,
     which also writes the 2-byte header and the combined-checksum trailer.
No codec arithmetic happens here; the compressor is `Lib.deflate_shard` (injectable so the
host-side logic can be exercised on CPU with gloo, where no GPU kernel can run).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

from . import binding as zb

CHUNK = 131072
WINDOW = 32768
ADLER_BASE = 65521


def shard_ranges(n: int, world: int, chunk: int = CHUNK) -> List[Tuple[int, int]]:
    """Contiguous chunk-aligned [begin, end) per rank, sizes differing by at most one chunk."""
    nchunks = (n + chunk - 1) // chunk
    out, c0 = [], 0
    for r in range(world):
        c1 = c0 + nchunks // world + (1 if r < nchunks % world else 0)
        out.append((min(n, c0 * chunk), min(n, c1 * chunk)))
        c0 = c1
    return out


def adler_join(a1: int, a2: int, len2: int) -> int:
    """Exact Adler-32 of A||B (the reference's adler32_combine keeps non-canonical 65521 folds)."""
    rem = len2 % ADLER_BASE
    s1, s2, t1, t2 = a1 & 0xFFFF, (a1 >> 16) & 0xFFFF, a2 & 0xFFFF, (a2 >> 16) & 0xFFFF
    return ((s1 + t1 + ADLER_BASE - 1) % ADLER_BASE) | (((s2 + t2 + rem * s1 + ADLER_BASE - rem) % ADLER_BASE) << 16)


@dataclass
class StreamPlan:
    offsets: List[int]          # byte offset of every rank's shard inside the final stream
    total: int                  # length of the final stream
    n_in: int                   # total uncompressed bytes
    crc32: int
    adler32: int
    header: bytes
    trailer: bytes


def zlib_header(level: int) -> bytes:
    level = 6 if level < 0 else level                                       # Z_DEFAULT_COMPRESSION, deflate.c:251
    fl = 0 if level < 2 else 1 if level < 6 else 2 if level == 6 else 3     # deflate.c:625-649
    h = (0x78 << 8) | (fl << 6)
    h += 31 - h % 31
    return bytes([h >> 8, h & 0xFF])


def plan_stream(metas: Sequence[Sequence[int]], level: int, wrap: int, crc_combine: Callable[[int, int, int], int]) -> StreamPlan:
    """metas[r] = (compressed bytes, input bytes, crc32, adler32) in rank order."""
    level = 6 if level < 0 else level
    header = zlib_header(level) if wrap == zb.WRAP_ZLIB else (
        bytes([31, 139, 8, 0, 0, 0, 0, 0, 2 if level == 9 else 4 if level < 2 else 0, 3]) if wrap == zb.WRAP_GZIP else b"")
    offsets, pos, crc, adl, n_in = [], len(header), 0, 1, 0
    for clen, n, c, a in metas:
        offsets.append(pos)
        pos += clen
        crc = crc_combine(crc, c, n) if n else crc
        adl = adler_join(adl, a, n)
        n_in += n
    if wrap == zb.WRAP_ZLIB:
        trailer = adl.to_bytes(4, "big")
    elif wrap == zb.WRAP_GZIP:
        trailer = crc.to_bytes(4, "little") + (n_in & 0xFFFFFFFF).to_bytes(4, "little")
    else:
        trailer = b""
    return StreamPlan(offsets, pos + len(trailer), n_in, crc, adl, header, trailer)


def exchange_meta(local: Sequence[int], device, group=None) -> List[List[int]]:
    """All-gather of the four per-rank integers (the path's one small collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor(list(local), dtype=torch.int64, device=device)
    out = torch.empty(world * 4, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.view(world, 4).cpu().tolist()


def _global(group, r: int) -> int:
    """send/recv take GLOBAL ranks; `root` and the loop indices here are ranks within `group`."""
    import torch.distributed as dist
    return r if group is None else dist.get_global_rank(group, r)


def gather_stream(local_bytes, local_len: int, plan: StreamPlan, root: int = 0, group=None):
    """Variable-length gather of the shards into one buffer on `root` (uint8 tensor) incl. header/trailer."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank != root:
        if local_len:
            dist.send(local_bytes[:local_len], dst=_global(group, root), group=group)
        return None
    out = torch.empty(plan.total, dtype=torch.uint8, device=local_bytes.device)
    if plan.header:
        out[:len(plan.header)] = torch.tensor(list(plan.header), dtype=torch.uint8, device=out.device)
    if plan.trailer:
        out[plan.total - len(plan.trailer):] = torch.tensor(list(plan.trailer), dtype=torch.uint8, device=out.device)
    ends = plan.offsets[1:] + [plan.total - len(plan.trailer)]
    for r in range(world):
        a, b = plan.offsets[r], ends[r]
        if r == root:
            out[a:b] = local_bytes[:b - a]
        elif b > a:
            dist.recv(out[a:b], src=_global(group, r), group=group)
    return out


def deflate_sharded(lib, data, halo, level: int = 6, wrap: int = zb.WRAP_ZLIB, group=None, out=None,
                    assemble: bool = True, root: int = 0, compress_fn: Optional[Callable] = None, stream=None):
    """Rank-local part of a multi-GPU deflate.

    data : this rank's contiguous slice of the input (uint8 tensor on this rank's device)
    halo : up to 32 KiB that precede it (uint8 tensor, same device) or None for rank 0
    Returns (plan, local_compressed_tensor, local_len, assembled_or_None).
    """
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = data.numel()
    last = rank == world - 1
    cap = n + (n >> 12) + (n >> 14) + 64
    if out is None:
        out = torch.empty(cap, dtype=torch.uint8, device=data.device)
    if compress_fn is None:
        flags = zb.ZB200_DEFLATE_NO_HEADER | zb.ZB200_DEFLATE_NO_TRAILER | (0 if last else zb.ZB200_DEFLATE_NOT_LAST)
        if halo is not None and halo.numel():
            joined = torch.cat([halo, data])                       # contiguous [dict][src] on the device
            dptr, sptr, dlen = joined.data_ptr(), joined.data_ptr() + halo.numel(), halo.numel()
        else:
            joined, dptr, sptr, dlen = data, 0, data.data_ptr(), 0
        clen, crc, adl = lib.deflate_shard(sptr, n, dptr if dlen else None, dlen, out.data_ptr(), out.numel(), level,
                                           zb.WRAP_RAW, flags, stream)
    else:
        clen, crc, adl = compress_fn(data, halo, last, out)
    metas = exchange_meta((clen, n, crc, adl), data.device, group)
    plan = plan_stream(metas, level, wrap, lib.crc32_combine)
    assembled = gather_stream(out, clen, plan, root, group) if assemble else None
    return plan, out, clen, assembled


# ---------------------------------------------------------------------------------------------
# One stream from all ranks with the gather hidden behind the compression (bench.py --gpus N)
# ---------------------------------------------------------------------------------------------
def piece_ranges(total: int, world: int, pieces_per_rank, chunk: int = CHUNK) -> List[List[Tuple[int, int]]]:
    """The input cut into rounds of `world` equal pieces dealt round robin: global piece g = j * world + r is round j of
    rank r.  pieces_per_rank is the number of (equal) rounds, or a tuple of fractions of a rank's share per round -- the
    LAST round's output is the only transfer that is not hidden behind compression, so it pays to make it the small one,
    e.g. (0.5, 0.375, 0.125).  Returns, per rank, its [begin, end) list in round order.  Pieces are chunk aligned (the
    last ones may be short or empty)."""
    fr = [1.0 / pieces_per_rank] * pieces_per_rank if isinstance(pieces_per_rank, int) else list(pieces_per_rank)
    nchunks = (total + chunk - 1) // chunk
    per_rank = (nchunks + world - 1) // world                      # chunks of a rank's share
    sizes, used = [], 0
    for k, f in enumerate(fr):
        c = per_rank - used if k == len(fr) - 1 else min(per_rank - used, max(1, round(per_rank * f)))
        sizes.append(c)
        used += c
    out: List[List[Tuple[int, int]]] = [[] for _ in range(world)]
    start = 0
    for c in sizes:
        for r in range(world):
            a = (start + r * c) * chunk
            out[r].append((min(total, a), min(total, a + c * chunk)))
        start += world * c
    return out


_SIDE = {}


def _side_stream(dev, which: int = 0):
    """Side streams per device for deflate_rounds: 0 = exchange and transfers, 1 = the odd rounds' kernels."""
    import torch
    key = (dev.type, dev.index, which)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def share_device_buffer(lib, nbytes: int, root: int = 0, group=None):
    """A device buffer on `root` that every rank of the box can name: the root allocates it (zb200_alloc_device), exports
    a CUDA IPC handle, the others open it.  Returns (pointer valid on this rank, closer) or (None, None) when IPC is not
    available here (then deflate_rounds falls back to NCCL send / recv)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    h = torch.zeros(65, dtype=torch.uint8, device=dev)
    ptr = None
    if rank == root:
        ptr = lib.dll.zb200_alloc_device(nbytes)
        buf = (C.c_ubyte * 64)()
        if ptr and lib.dll.zb200_ipc_export(C.c_void_p(ptr), buf) == 0:
            h[:64] = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
            h[64] = 1
    dist.broadcast(h, src=_global(group, root), group=group)
    ok = int(h[64]) == 1
    if ok and rank != root:
        buf = (C.c_ubyte * 64)(*h[:64].cpu().tolist())
        p = C.c_void_p(0)
        ok = lib.dll.zb200_ipc_open(buf, C.byref(p)) == 0 and bool(p.value)
        ptr = p.value if ok else None
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if not bool(flag.item()):
        if ptr and rank == root:
            lib.dll.zb200_free_device(C.c_void_p(ptr))
        elif ptr:
            lib.dll.zb200_ipc_close(C.c_void_p(ptr))
        return None, None

    def close():
        dist.barrier(group)
        if rank == root:
            lib.dll.zb200_free_device(C.c_void_p(ptr))
        else:
            lib.dll.zb200_ipc_close(C.c_void_p(ptr))
    return ptr, close


def deflate_rounds(lib, pieces, level: int = 1, wrap: int = zb.WRAP_ZLIB, group=None, root: int = 0, outs=None,
                   final=None, host_final: Optional[int] = None, stream=None, copy_stream=None, last_round_is_last: bool = True,
                   peer_final: Optional[int] = None, placed: Optional[list] = None):
    """ONE stream from all ranks, assembled on `root` while the compression is still running.

    pieces[j] = (src_ptr, n, dict_ptr, dict_len) for round j: this rank's j-th piece (device or pinned host memory) and
    the <= 32 KiB in front of it.  The stream is the pieces in GLOBAL order g = j * world + r, so after round j
    every rank knows (one all-gather of {len, n, crc32, adler32}) where its round-j output lands in the final stream,
    and the transfer of round j -- NCCL send/recv into `final` on the root (device mode), or a D2H copy straight into
    the shared, page-locked host buffer at `host_final` (host mode) -- runs while round j + 1 is compressed.  Only the
    last round's transfer is exposed.  With `peer_final` (the root's buffer opened through CUDA IPC, share_device_buffer)
    the device-mode transfers are plain copies issued by their OWNER on its copy stream: the copy engines move them over
    NVLink, no SM on either side is needed (NCCL's send / recv kernels wait for SM slots that the persistent walk CTAs hold).
    `placed`, if given, receives (offset, length) of this rank's part of every round.  The root adds the header and the trailer with the combined checksum.
    outs[j]: this rank's device output buffer of round j.  Returns (total stream length, crc32, adler32, n_in).
    """
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lvl = 6 if level < 0 else level
    header = zlib_header(lvl) if wrap == zb.WRAP_ZLIB else (
        bytes([31, 139, 8, 0, 0, 0, 0, 0, 2 if lvl == 9 else 4 if lvl < 2 else 0, 3]) if wrap == zb.WRAP_GZIP else b"")
    rounds = len(pieces)
    pos, crc_all, adl_all, n_all = len(header), 0, 1, 0
    works = []
    dev = outs[0].device
    # Pipelined when the library offers the two halves of the call: piece j + 1 is enqueued before piece j's length is read,
    # and the exchange and the transfers run on a side stream, so nothing of them queues behind the next piece's kernels.
    pipelined = hasattr(lib, "deflate_shard_begin") and dev.type == "cuda"
    side = _side_stream(dev) if pipelined else None
    # (running odd rounds on a second compute stream was tried: 14.75 -> 14.64 ms device resident at N = 2, but the host-buffer
    # run lost 4 ms -- the two pieces' H2D copies then share the link and the first piece's D2H starts later)
    import contextlib

    def on_side():
        return torch.cuda.stream(side) if side is not None else contextlib.nullcontext()

    def args_of(j):
        src_ptr, n, dict_ptr, dict_len = pieces[j]
        last = last_round_is_last and j == rounds - 1 and rank == world - 1
        flags = zb.ZB200_DEFLATE_NO_HEADER | zb.ZB200_DEFLATE_NO_TRAILER | (0 if last else zb.ZB200_DEFLATE_NOT_LAST)
        return (src_ptr, n, dict_ptr if dict_len else None, dict_len, outs[j].data_ptr(), outs[j].numel(), level, zb.WRAP_RAW, flags, stream)

    jobs = {}
    if pipelined and rounds:
        jobs[0] = lib.deflate_shard_begin(*args_of(0))
    for j in range(rounds):
        n = pieces[j][1]
        if pipelined:
            if j + 1 < rounds:
                jobs[j + 1] = lib.deflate_shard_begin(*args_of(j + 1))
            clen, crc, adl = lib.deflate_shard_end(jobs.pop(j))
        else:
            clen, crc, adl = lib.deflate_shard(*args_of(j))
        with on_side():
            metas = exchange_meta((clen, n, crc, adl), dev, group)
            offs = []
            for cl, m, c, a in metas:                              # global order within a round = rank order
                offs.append(pos)
                pos += cl
                if m:
                    crc_all = lib.crc32_combine(crc_all, c, m)
                    adl_all = adler_join(adl_all, a, m)
                n_all += m
            if placed is not None:
                placed.append((offs[rank], clen))
            if host_final is not None or peer_final is not None:   # every rank writes its own part of the shared buffer
                if clen:
                    base = host_final if host_final is not None else peer_final
                    lib._check(lib.dll.zb200_copy_async(base + offs[rank], outs[j].data_ptr(), clen, zb._stream(copy_stream)),
                               "zb200_copy_async")
            elif final is not None or rank != root:
                ops = []
                if rank == root:
                    for r in range(world):
                        cl = metas[r][0]
                        if not cl:
                            continue
                        if r == root:
                            final[offs[r]:offs[r] + cl].copy_(outs[j][:cl], non_blocking=True)
                        else:
                            ops.append(dist.P2POp(dist.irecv, final[offs[r]:offs[r] + cl], _global(group, r), group))
                elif clen:
                    ops.append(dist.P2POp(dist.isend, outs[j][:clen], _global(group, root), group))
                if ops:
                    works += dist.batch_isend_irecv(ops)
    if wrap == zb.WRAP_ZLIB:
        trailer = adl_all.to_bytes(4, "big")
    elif wrap == zb.WRAP_GZIP:
        trailer = crc_all.to_bytes(4, "little") + (n_all & 0xFFFFFFFF).to_bytes(4, "little")
    else:
        trailer = b""
    total = pos + len(trailer)
    with on_side():
        for w in works:
            w.wait()
        if rank == root:
            if host_final is not None:
                import ctypes as C
                C.memmove(host_final, header, len(header))
                C.memmove(host_final + pos, trailer, len(trailer))
            elif peer_final is not None:
                ht = torch.tensor(list(header + trailer), dtype=torch.uint8, device=dev)
                if header:
                    lib._check(lib.dll.zb200_copy_async(peer_final, ht.data_ptr(), len(header), zb._stream(copy_stream)), "zb200_copy_async")
                if trailer:
                    lib._check(lib.dll.zb200_copy_async(peer_final + pos, ht.data_ptr() + len(header), len(trailer), zb._stream(copy_stream)), "zb200_copy_async")
            elif final is not None:
                if header:
                    final[:len(header)] = torch.tensor(list(header), dtype=torch.uint8, device=dev)
                if trailer:
                    final[pos:total] = torch.tensor(list(trailer), dtype=torch.uint8, device=dev)
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)              # the caller's stream sees the assembled stream
    if host_final is not None or peer_final is not None:
        lib._check(lib.dll.zb200_sync(zb._stream(copy_stream)), "zb200_sync")
    if peer_final is not None:
        dist.barrier(group)                                        # every rank's copies have landed: the stream is complete on the root
    return total, crc_all, adl_all, n_all


# ---------------------------------------------------------------------------------------------
# BASELINE config 4: crc32 / adler32 of one buffer spread over the ranks
# ---------------------------------------------------------------------------------------------
def checksum_sharded(lib, data, group=None, checksum_fn: Optional[Callable] = None, stream=None) -> Tuple[int, int, int]:
    """crc32 / adler32 of the concatenation of every rank's slice, in rank order.

    Each rank runs the fused checksum kernel over its own slice; the exchange is one all-gather of
    {crc32, adler32, length} per rank, and every rank folds the triples in rank order with
    crc32_combine (qcsrc/crc32.c:370) and the Adler join -- the combine is not commutative, so a
    reduction collective cannot do it.  Returns (crc32, adler32, total length).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = data.numel()
    if checksum_fn is None:
        crc, adl = lib.checksum(data.data_ptr(), n, stream) if n else (0, 1)
    else:
        crc, adl = checksum_fn(data)
    mine = torch.tensor([crc, adl, n], dtype=torch.int64, device=data.device)
    out = torch.empty(world * 3, dtype=torch.int64, device=data.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    crc_all, adl_all, total = 0, 1, 0
    for c, a, m in out.view(world, 3).cpu().tolist():
        if m:
            crc_all = lib.crc32_combine(crc_all, c, m)
            adl_all = adler_join(adl_all, a, m)
        total += m
    return crc_all, adl_all, total


# ---------------------------------------------------------------------------------------------
# BASELINE config 3: batch inflate of independent streams, sharded by stream
# ---------------------------------------------------------------------------------------------
def balance_streams(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Contiguous ranges of stream indices per rank with nearly equal compressed bytes (streams stay in order, so
    outputs concatenate in rank order).  Returns world lists of indices."""
    total = sum(sizes)
    out, i, acc = [], 0, 0
    for r in range(world):
        want = total * (r + 1) / world
        j = i
        while j < len(sizes) and (r == world - 1 or acc + sizes[j] / 2 <= want):
            acc += sizes[j]
            j += 1
        out.append(list(range(i, j)))
        i = j
    return out


def inflate_sharded(lib, streams: Sequence[bytes], caps: Sequence[int], wrap: int = zb.WRAP_ZLIB, group=None,
                    inflate_fn: Optional[Callable] = None):
    """Every rank decodes its share of the streams; nothing but {status, length} per stream is exchanged.

    Returns (my_indices, my_outputs, statuses_of_all_streams, lengths_of_all_streams)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = balance_streams([len(z) for z in streams], world)
    mine = parts[rank]
    fn = inflate_fn or (lambda zs, cs: lib.inflate_batch(zs, cs, wrap))
    outs, st = fn([streams[i] for i in mine], [caps[i] for i in mine]) if mine else ([], [])
    n = len(streams)
    local = torch.full((n, 2), -100, dtype=torch.int64)          # below every zlib status (>= -6) and every length
    for k, i in enumerate(mine):
        local[i, 0], local[i, 1] = st[k], len(outs[k])
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    local = local.to(dev)
    dist.all_reduce(local, op=dist.ReduceOp.MAX, group=group)     # every stream has exactly one owner; the rest hold -100
    res = local.cpu().tolist()
    return mine, outs, [int(r[0]) for r in res], [int(r[1]) for r in res]


def inflate_sharded_dev(lib, d_src, src_off, d_dst, dst_off, wrap: int = zb.WRAP_ZLIB, group=None, stream=None):
    """Device-resident form of inflate_sharded for the timed runs: every rank holds the stream set (d_src with host
    offsets src_off[n + 1]) and the output arena; rank r decodes its contiguous, size-balanced range into its own part of
    d_dst.  Exchange: one all-reduce(MAX) that fills in {status, length} of every stream on every rank.
    Returns (first, last) stream index of this rank and the [n, 2] tensor of {status, length}."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(src_off) - 1
    sizes = [int(src_off[i + 1] - src_off[i]) for i in range(n)]
    parts = balance_streams(sizes, world)
    mine = parts[rank]
    dev = d_src.device
    res = torch.full((n, 2), -100, dtype=torch.int64, device=dev)
    a, b = (mine[0], mine[-1] + 1) if mine else (0, 0)
    if b > a:
        d_so = torch.as_tensor(src_off[a:b + 1], dtype=torch.int64).to(dev)
        d_do = torch.as_tensor(dst_off[a:b + 1], dtype=torch.int64).to(dev)
        d_len = torch.zeros(b - a, dtype=torch.int64, device=dev)
        d_st = torch.zeros(b - a, dtype=torch.int32, device=dev)
        lib.inflate_batch_dev(d_src.data_ptr(), d_so.data_ptr(), b - a, d_dst.data_ptr(), d_do.data_ptr(), d_len.data_ptr(),
                              d_st.data_ptr(), wrap, stream)
        res[a:b, 0] = d_st.to(torch.int64)
        res[a:b, 1] = d_len
    dist.all_reduce(res, op=dist.ReduceOp.MAX, group=group)
    return (a, b), res


# ---------------------------------------------------------------------------------------------
# BASELINE config 5: one ZIP archive whose members are compressed on all ranks
# ---------------------------------------------------------------------------------------------
def assign_files(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first: files by decreasing size, each to the rank with the least bytes so far
    (ties to the lowest rank).  Every rank's list is returned in increasing index order."""
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda k: (-sizes[k], k)):
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(i)
        load[r] += sizes[i]
    return [sorted(x) for x in out]


def zip_sharded(lib, names: Sequence[str], datas: Sequence[bytes], level: int = 6, group=None, root: int = 0,
                segment_fn: Optional[Callable] = None, directory_fn: Optional[Callable] = None):
    """One ZIP32 archive from files compressed on every rank.

    The unit is the file (independent raw-deflate streams): files are dealt to the ranks by size, every rank turns
    its share into a *segment* -- [local header | name | data] records laid out on its GPU by zb200_zip_segment --
    and the archive is the segments in rank order followed by the central directory.  Exchange: one all-gather of
    the per-member records {file index, local offset, compressed size, raw size, crc32} (padded to the largest share)
    and the variable-length gather of the segments to the root, which appends the directory (zb200_zip_directory).
    Returns the archive (bytes) on the root, None elsewhere.
    """
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shares = assign_files([len(d) for d in datas], world)
    mine = shares[rank]
    seg_fn = segment_fn or (lambda nm, ds: lib.zip_segment(nm, ds, level))
    dir_fn = directory_fn or (lambda nm, metas, cd_off: lib.zip_directory(nm, metas, cd_off))
    seg, metas = seg_fn([names[i] for i in mine], [datas[i] for i in mine]) if mine else (b"", [])
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    width = max(len(x) for x in shares)
    rec = torch.full((width + 1, 5), -1, dtype=torch.int64)
    rec[0, 0] = len(seg)
    for k, (i, m) in enumerate(zip(mine, metas)):
        rec[k + 1] = torch.tensor([i, m[0], m[1], m[2], m[3]], dtype=torch.int64)
    rec = rec.to(dev).view(-1)
    allrec = torch.empty(world * (width + 1) * 5, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allrec, rec, group=group)
    allrec = allrec.view(world, width + 1, 5).cpu().tolist()
    seg_len = [int(allrec[r][0][0]) for r in range(world)]
    seg_t = torch.frombuffer(bytearray(seg), dtype=torch.uint8).to(dev) if seg else torch.empty(0, dtype=torch.uint8, device=dev)
    if rank != root:
        if len(seg):
            dist.send(seg_t, dst=_global(group, root), group=group)
        return None
    parts, base, order, glob = [], 0, [], []
    for r in range(world):
        if r == root:
            parts.append(seg)
        elif seg_len[r]:
            buf = torch.empty(seg_len[r], dtype=torch.uint8, device=dev)
            dist.recv(buf, src=_global(group, r), group=group)
            parts.append(bytes(buf.cpu().numpy()))
        else:
            parts.append(b"")
        for k in range(len(shares[r])):
            i, lo, cl, rl, crc = allrec[r][k + 1]
            order.append(names[i])
            glob.append((base + lo, cl, rl, crc))
        base += seg_len[r]
    return b"".join(parts) + dir_fn(order, glob, base)
