"""Multi-GPU check of the sharded paths (run under torchrun): deflate assembly, checksum combine, batch inflate."""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from zlib_b200 import load, binding as zb, dist as zd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
L = load()
assert L.dll.zb200_init(local) == 0, L.last_error()
n = 64 << 20
data = L.synth(n, kind=1, seed=77)                       # the same buffer on every rank
a, b = zd.shard_ranges(n, world)[rank]
mine = torch.from_numpy(data[a:b].copy()).cuda()
# config 4
crc, adl, total = zd.checksum_sharded(L, mine)
assert (crc, adl, total) == (zlib.crc32(data.tobytes()), zlib.adler32(data.tobytes()), n), (crc, adl, total)
# config 2 across ranks
halo = torch.from_numpy(data[max(0, a - zd.WINDOW):a].copy()).cuda() if a > 0 else None
plan, out, clen, full = zd.deflate_sharded(L, mine, halo, 1, zb.WRAP_ZLIB)
if rank == 0:
    assert zlib.decompress(bytes(full.cpu().numpy())) == data.tobytes()
# config 3
sz = 65536
zs = [zlib.compress(data[i * sz:(i + 1) * sz].tobytes(), 6) for i in range(512)]
idx, outs, st, lens = zd.inflate_sharded(L, zs, [sz] * len(zs))
assert st == [0] * len(zs) and lens == [sz] * len(zs)
assert all(outs[k] == data[i * sz:(i + 1) * sz].tobytes() for k, i in enumerate(idx))
# config 5: files dealt to the ranks, one archive on the root
import io, random, zipfile
rng = random.Random(5)
fsz = [int(4096 * 2 ** rng.uniform(0, 10)) for _ in range(300)]
foff = [0]
for z in fsz:
    foff.append(foff[-1] + z)
blob = L.synth(foff[-1], kind=1, seed=78)
names = [f"f{i:05d}.bin" for i in range(len(fsz))]
datas = [blob[foff[i]:foff[i + 1]].tobytes() for i in range(len(fsz))]
arc = zd.zip_sharded(L, names, datas, 1)
if rank == 0:
    zf = zipfile.ZipFile(io.BytesIO(arc))
    assert zf.testzip() is None and sorted(zf.namelist()) == names
    assert all(zf.read(nm) == d for nm, d in zip(names, datas))
dist.barrier()
if rank == 0:
    print(f"dist_check ok on {world} GPUs: crc {crc:08x}, stream {plan.total} bytes, {len(zs)} streams inflated")
dist.destroy_process_group()
