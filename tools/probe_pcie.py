"""Copy-only ceiling of the host-buffer path at N ranks (run directly or under torchrun): every rank moves 1 GiB H2D and
0.5 GiB D2H between pinned host memory and its GPU -- alone, both at once -- and all ranks do so at the same time.
Prints one JSON line per rank 0: per-rank and aggregate GB/s, and where the pinned pages and the GPUs sit (NUMA)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
h_in.fill_(1); h_out.fill_(2)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def t(fn, reps=5):
    fn(); barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    barrier()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        x = torch.tensor([dt], device="cuda"); dist.all_reduce(x, op=dist.ReduceOp.MAX); dt = float(x.item())
    return dt

def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()

a, b, c = t(h2d), t(d2h), t(both)
numa = {}
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    pci = pynvml.nvmlDeviceGetPciInfo(h).busId
    pci = pci.decode() if isinstance(pci, bytes) else pci
    p = f"/sys/bus/pci/devices/{pci.lower()[-12:]}/numa_node"
    numa["gpu_numa_node"] = open(p).read().strip() if os.path.exists(p) else "?"
except Exception as e:   # noqa: BLE001
    numa["gpu_numa_node"] = f"? ({e})"
try:
    numa["host_numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
except Exception:   # noqa: BLE001
    numa["host_numa_nodes"] = "?"
numa["cpus_allowed"] = len(os.sched_getaffinity(0))
if world > 1:
    alln = [None] * world
    dist.all_gather_object(alln, numa)
else:
    alln = [numa]
if rank == 0:
    print(json.dumps({"n_ranks": world, "h2d_1GiB_ms": round(a * 1e3, 2), "h2d_GBps_per_rank": round(n / a / 1e9, 1), "h2d_GBps_aggregate": round(world * n / a / 1e9, 1),
                      "d2h_0.5GiB_ms": round(b * 1e3, 2), "d2h_GBps_per_rank": round(n / 2 / b / 1e9, 1),
                      "both_ms": round(c * 1e3, 2), "both_GBps_aggregate_in_plus_out": round(world * 1.5 * n / c / 1e9, 1),
                      "copy_only_ceiling_GBps_of_uncompressed_input": round(world * n / c / 1e9, 1), "numa": alln}))
if world > 1:
    dist.destroy_process_group()
