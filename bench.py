#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark (one JSON line on stdout).

Workload = BASELINE.json configs[1]: 1 GiB of the synthetic mixed corpus,
deflate in 128 KiB chunks on one B200, level 1 as the headline (`value`), with level 6, batch
inflate and the fused checksum reported beside it under `extra`.

  value     deflate level-1 throughput, GB/s of UNCOMPRESSED bytes, input and output resident in HBM,
            timed with CUDA events on the launching stream (max over ranks for --gpus N).
  e2e       the same metric through the reference-facing call compress2(dest, &destLen, source,
            sourceLen, 1) of libzb200.so with HOST (pinned) buffers: H2D + kernels + D2H in the timed region.
  roofline  dominant kernel of a step (per-kernel CUDA events inside the library, zb200_profile):
            algorithmic bytes (n_in + n_out) / that kernel's time, against the measured HBM peak.
  cpu_baseline / --impl reference
            the UNMODIFIED reference (oracle/_ref/libzref.so, built from /root/reference by oracle/Makefile)
            running compress2 level 1 on the host cores of this box, all threads, on a bounded sample.

A step = one pass of the hot path over the 1 GiB batch (per GPU: weak scaling).
"""
import argparse
import ctypes as C
import json
import re
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "deflate_level1_GBps_uncompressed"
UNIT = "GB/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:   # noqa: BLE001
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md).  NVML when importable (a sample
    every 2 ms: the timed region is ~0.1 s), else the nvidia-smi query of the profiling recipe."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:   # noqa: BLE001
            self.nvml = None

    def _nvml_row(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = [bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)]
        return [str(mhz), str(self.max_mhz)] + ["Active" if f else "Not Active" for f in flags]

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._nvml_row())
                else:
                    o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                       capture_output=True, text=True, timeout=5).stdout.strip()
                    if o:
                        self.rows.append([x.strip() for x in o.split(",")])
            except Exception:   # noqa: BLE001
                pass
            self.stop_flag.wait(0.002 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for i, nme in enumerate(self.NAMES):
                if len(r) > 2 + i and r[2 + i].lower().startswith("active"):
                    reasons.add(nme)
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------- reference arm
def reference_compress_throughput(level, sample_bytes, kind=1, seed=1):
    """compress2 of the unmodified reference on all host cores; returns (GB/s, cores, seconds, ratio)."""
    import zhelpers
    from zlib_b200 import load
    if not os.path.exists(zhelpers.REF_PATH):
        return None
    ref = zhelpers.Ref()
    lib = load()
    cores = os.cpu_count() or 1
    per = max(131072, (sample_bytes // cores) // 131072 * 131072)
    data = lib.synth(per * cores, kind=kind, seed=seed)          # host generator only: no GPU involved
    outs = [0] * cores
    cap = ref.dll.compressBound(per)
    bufs = [C.create_string_buffer(cap) for _ in range(cores)]

    def work(i):
        ol = C.c_ulong(cap)
        rc = ref.dll.compress2(bufs[i], C.byref(ol), C.c_void_p(data.ctypes.data + i * per), per, level)
        assert rc == 0
        outs[i] = ol.value

    ths = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return per * cores / dt / 1e9, cores, dt, per * cores / max(1, sum(outs)), per * cores


def reference_side_paths(sample_per_thread=16 << 20, kind=1, seed=1):
    """The other paths of the reference on all host cores, bounded samples (reported beside the GPU extras):
    compress2 level 6, uncompress of those streams, crc32 and adler32 -- every thread on its own slice."""
    import zhelpers
    from zlib_b200 import load
    if not os.path.exists(zhelpers.REF_PATH):
        return None
    ref = zhelpers.Ref()
    lib = load()
    cores = os.cpu_count() or 1
    per = sample_per_thread
    data = lib.synth(per * cores, kind=kind, seed=seed)
    cap = ref.dll.compressBound(per)
    zbufs = [C.create_string_buffer(cap) for _ in range(cores)]
    obufs = [C.create_string_buffer(per) for _ in range(cores)]
    zlen = [0] * cores
    sink = [0] * cores

    def run(fn):
        ths = [threading.Thread(target=fn, args=(i,)) for i in range(cores)]
        t0 = time.perf_counter()
        [t.start() for t in ths]
        [t.join() for t in ths]
        return time.perf_counter() - t0

    def f_comp(i):
        ol = C.c_ulong(cap)
        assert ref.dll.compress2(zbufs[i], C.byref(ol), C.c_void_p(data.ctypes.data + i * per), per, 6) == 0
        zlen[i] = ol.value

    def f_unc(i):
        ol = C.c_ulong(per)
        assert ref.dll.uncompress(obufs[i], C.byref(ol), zbufs[i], zlen[i]) == 0 and ol.value == per

    def f_crc(i):
        for _ in range(8):
            sink[i] = ref.dll.crc32(0, C.c_void_p(data.ctypes.data + i * per), per)

    def f_adl(i):
        for _ in range(8):
            sink[i] = ref.dll.adler32(1, C.c_void_p(data.ctypes.data + i * per), per)

    tot = per * cores
    out = {"cores": cores, "sample": f"{per >> 20} MiB of the mixed corpus per host thread"}
    out["compress2_level6_GBps"] = round(tot / run(f_comp) / 1e9, 4)
    out["level6_ratio"] = round(tot / max(1, sum(zlen)), 4)
    out["uncompress_GBps"] = round(tot / run(f_unc) / 1e9, 4)
    out["crc32_GBps"] = round(8 * tot / run(f_crc) / 1e9, 3)
    out["adler32_GBps"] = round(8 * tot / run(f_adl) / 1e9, 3)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    size = args.size_mib << 20
    cores = os.cpu_count() or 1
    sample = min(size, cores * (48 << 20))
    vals, ratio, nbytes = [], None, None
    for i in range(args.warmup + args.steps):
        r = reference_compress_throughput(1, sample)
        if r is None:
            emit_line({"impl": "reference", "unavailable": "oracle/_ref/libzref.so not built on this box"})
            return 0
        if i >= args.warmup:
            vals.append(r[0])
        ratio, nbytes, cores = r[3], r[4], r[1]
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(nbytes / v / 1e6, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.size_mib} MiB synthetic mixed corpus, deflate level 1, 128 KiB chunks",
                       "level": 1, "sample_bytes": nbytes},
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"compress2 level 1 over {nbytes >> 20} MiB of the mixed corpus, one {nbytes // cores >> 20} MiB slice per host thread"},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ratio": round(ratio, 4), "gpu_launches": 0}
    emit_line(line)
    return 0


# ------------------------------------------------------------------------------------- our arm
_REAL_STDOUT = None


def claim_stdout():
    """stdout carries ONE JSON line.  Libraries write there too (NCCL prints its version banner on the first communicator),
    so file descriptor 1 is pointed at stderr for the run and the line goes to a duplicate of the original."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zb200", choices=["zb200", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024)
    ap.add_argument("--no-extra", action="store_true", help="skip level 6 / inflate / checksum side measurements")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from zlib_b200 import load, binding as zb, dist as zdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load()
    assert lib.dll.zb200_init(local) == 0, lib.last_error()
    s = torch.cuda.current_stream()

    n = args.size_mib << 20
    warm = max(3, args.warmup)
    log(f"[rank {rank}] generating {args.size_mib} MiB mixed corpus")
    pin_src = lib.dll.zb200_alloc_pinned(n)
    cap = lib.compress_bound(n) + 64
    pin_dst = lib.dll.zb200_alloc_pinned(cap)
    assert pin_src and pin_dst
    lib.dll.zb200_synth(C.c_void_p(pin_src), n, 1, 1 + rank)
    host = np.ctypeslib.as_array(C.cast(pin_src, C.POINTER(C.c_uint8)), shape=(n,))
    d_src = torch.empty(n, dtype=torch.uint8, device=dev)
    lib.dll.zb200_copy(C.c_void_p(d_src.data_ptr()), C.c_void_p(pin_src), n, None)
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    halo = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {}

    def step_deflate(level):
        if world == 1:
            state["clen"] = lib.deflate(d_src.data_ptr(), n, d_dst.data_ptr(), cap, level, zb.WRAP_ZLIB, s)
        else:   # shard + all-gather of {len, n, crc, adler}; bytes stay where they are (offsets known to all)
            plan, _, clen, _ = zdist.deflate_sharded(lib, d_src, halo, level, zb.WRAP_ZLIB, out=d_dst, assemble=False, stream=s)
            state["clen"], state["plan"] = clen, plan

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.kernel_launches()
        e0.record(s)
        for _ in range(steps):
            fn()
        e1.record(s)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, lib.kernel_launches() - l0

    # ---- headline: deflate level 1, device resident ----
    sampler = ClockSampler(local)
    sampler.start()
    ms1, launches = timed(lambda: step_deflate(1), args.steps, warm)
    clocks = sampler.summary()
    clen1 = state["clen"]
    value = world * n / ms1 / 1e6
    log(f"[rank {rank}] deflate L1: {ms1:.2f} ms/step  {value:.2f} GB/s aggregate  ratio {n / clen1:.3f}")

    # ---- per-kernel times (separate pass; events add overhead, so not part of `value`) ----
    lib.profile(True)
    step_deflate(1)
    prof = lib.profile_report()
    lib.profile(False)
    def kname(k):                                        # "(k_lz_walk<kWalkThreadsFast, false>)" -> "k_lz_walk"
        return re.sub(r"<.*>|[()]|zb::", "", k).strip()
    prof = {kname(k): v for k, v in prof.items()}
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else ("none", (0.0, 1))
    dom_launches = max(1, dom[1][1])                    # one launch per pipeline slab
    dom_ms = dom[1][0] / dom_launches                   # average launch duration of the dominant kernel
    peaks, peak_kind = measured_peaks()
    algo_bytes = n + clen1                              # every input byte read once, every output byte written once
    launch_bytes = algo_bytes / dom_launches            # algorithmic bytes one launch (one slab) accounts for
    achieved = launch_bytes / dom_ms / 1e6 if dom_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = {kname(k): v for k, v in json.load(open(tpath)).items()}.get(dom[0])
        except Exception:   # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": round(achieved, 2), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(achieved / peaks["hbm_gbs"], 5), "traffic": traffic, "peak_source": peak_kind,
                "algorithmic_bytes": algo_bytes, "algorithmic_bytes_per_launch": round(launch_bytes),
                "launches_per_step": dom_launches, "kernel_ms": round(dom_ms, 4),
                "kernel_share_of_step": round(dom[1][0] / sum(v[0] for v in prof.values()), 4) if prof else None,
                "kernels_ms": {k.replace("zb::", ""): round(v[0], 4) for k, v in prof.items()}}

    # ---- end to end through compress2 with host buffers (H2D + kernels + D2H inside the timed region) ----
    def step_e2e():
        ol = C.c_ulong(cap)
        rc = lib.dll.compress2(C.c_void_p(pin_dst), C.byref(ol), C.c_void_p(pin_src), n, 1)
        assert rc == 0, (rc, lib.last_error())
        state["e2e_len"] = ol.value

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": round(world * n / e2e_s / 1e9, 4), "unit": UNIT, "h2d_bytes_per_step": n,
           "d2h_bytes_per_step": int(state["e2e_len"]) + 32, "api": "compress2(dest,&destLen,source,sourceLen,1) on pinned host buffers"}
    log(f"[rank {rank}] e2e compress2: {e2e_s * 1e3:.1f} ms/step {e2e['value']} GB/s")

    extra, verified, cpu = {}, None, None
    if rank == 0:
        import zhelpers
        # ---- correctness gate (a) on the measured output: the reference decodes it bit-exact ----
        if not args.no_verify and world == 1:
            comp = np.ctypeslib.as_array(C.cast(pin_dst, C.POINTER(C.c_uint8)), shape=(int(state["e2e_len"]),))
            t0 = time.perf_counter()
            if os.path.exists(zhelpers.REF_PATH):
                ref = zhelpers.Ref()
                outb = np.empty(n, dtype=np.uint8)
                ol = C.c_ulong(n)
                rc = ref.dll.uncompress(C.c_void_p(outb.ctypes.data), C.byref(ol), C.c_void_p(comp.ctypes.data), len(comp))
                verified = bool(rc == 0 and ol.value == n and np.array_equal(outb, host))
                who = "reference uncompress (oracle/_ref)"
            else:
                orc = zhelpers.Oracle()
                rc, outb, used = orc.inflate(comp, n)
                verified = bool(rc == 0 and outb == host.tobytes())
                who = "oracle port"
            log(f"[rank 0] output verified by {who}: {verified} ({time.perf_counter() - t0:.1f} s)")
            assert verified, "reference could not decode the GPU deflate output"
        # ---- CPU baseline beside it ----
        cores = os.cpu_count() or 1
        r = reference_compress_throughput(1, min(n, cores * (48 << 20)))
        if r is not None:
            cpu = {"value": round(r[0], 4), "unit": UNIT, "cores": r[1], "kind": "reference",
                   "sample": f"reference compress2 level 1 over {r[4] >> 20} MiB of the same corpus, one {r[4] // r[1] >> 20} MiB slice per host thread",
                   "ratio": round(r[3], 4)}
            log(f"[rank 0] reference compress2 L1 on {r[1]} host threads: {r[0]:.3f} GB/s ratio {r[3]:.3f}")

    if not args.no_extra and world == 1:
        ms6, _ = timed(lambda: step_deflate(6), max(1, args.steps // 2), 1)
        extra["deflate_level6"] = {"GBps": round(n / ms6 / 1e6, 3), "ratio": round(n / state["clen"], 4), "ms": round(ms6, 2)}
        out2 = torch.zeros(2, dtype=torch.int32, device=dev)
        msc, _ = timed(lambda: lib.checksum_dev(d_src.data_ptr(), n, out2.data_ptr(), s), 10, 3)
        extra["crc32_adler32_fused"] = {"GBps": round(n / msc / 1e6, 1), "ms": round(msc, 4),
                                        "frac_of_hbm_peak": round(n / msc / 1e6 / peaks["hbm_gbs"], 4)}
        # batch inflate: 64 KiB zlib streams of the same corpus, level 6, produced by the reference when available
        import zhelpers
        sz, distinct = 65536, 2048
        ns = n // sz
        t0 = time.perf_counter()
        if os.path.exists(zhelpers.REF_PATH):
            ref = zhelpers.Ref()
            zs, src_kind = [None] * distinct, "reference compress2 level 6"

            def mk(lo, hi):
                for i in range(lo, hi):
                    zs[i] = ref.compress2(host[i * sz:(i + 1) * sz], 6)
            nt = min(os.cpu_count() or 1, 32)
            ths = [threading.Thread(target=mk, args=(distinct * t // nt, distinct * (t + 1) // nt)) for t in range(nt)]
            [t.start() for t in ths]
            [t.join() for t in ths]
        else:
            import zlib as pyz
            zs, src_kind = [pyz.compress(host[i * sz:(i + 1) * sz].tobytes(), 6) for i in range(distinct)], "system zlib level 6"
        zs = (zs * ((ns + distinct - 1) // distinct))[:ns]
        src_off = np.zeros(ns + 1, dtype=np.int64)
        src_off[1:] = np.cumsum([len(z) for z in zs])
        d_z = torch.from_numpy(np.frombuffer(b"".join(zs) + b"\0" * 8, dtype=np.uint8).copy()).to(dev)
        d_so = torch.from_numpy(src_off).to(dev)
        d_do = torch.arange(ns + 1, dtype=torch.int64, device=dev) * sz
        d_len = torch.zeros(ns, dtype=torch.int64, device=dev)
        d_st = torch.zeros(ns, dtype=torch.int32, device=dev)
        log(f"[rank 0] prepared {ns} streams ({src_kind}) in {time.perf_counter() - t0:.1f} s")
        msi, _ = timed(lambda: lib.inflate_batch_dev(d_z.data_ptr(), d_so.data_ptr(), ns, d_src.data_ptr(), d_do.data_ptr(),
                                                     d_len.data_ptr(), d_st.data_ptr(), zb.WRAP_ZLIB, s), 3, 2)
        ok = int(d_st.abs().sum()) == 0 and bool((d_len == sz).all())
        extra["inflate_batch"] = {"GBps": round(ns * sz / msi / 1e6, 3), "streams": ns, "stream_bytes": sz, "ms": round(msi, 2),
                                  "source": src_kind, "all_ok": ok, "compressed_fraction": round(int(src_off[-1]) / (ns * sz), 4)}
        # the same batch through zb200_inflate_batch with pinned HOST arenas: H2D + decode + D2H inside the timed region
        zin = int(src_off[-1])
        pin_z = lib.dll.zb200_alloc_pinned(zin + 64)
        assert pin_z
        z_host = d_z.cpu().numpy()                              # keep the array alive across the copy
        C.memmove(C.c_void_p(pin_z), C.c_void_p(z_host.ctypes.data), zin)
        del z_host
        h_so, h_do = src_off.astype(np.uint64), (np.arange(ns + 1, dtype=np.uint64) * sz)
        h_len, h_st = np.zeros(ns, dtype=np.uint64), np.zeros(ns, dtype=np.int32)
        th = []
        for _ in range(3):
            t0 = time.perf_counter()
            rc = lib.dll.zb200_inflate_batch(C.c_void_p(pin_z), C.c_void_p(h_so.ctypes.data), ns, C.c_void_p(pin_dst), C.c_void_p(h_do.ctypes.data),
                                             C.c_void_p(h_len.ctypes.data), C.c_void_p(h_st.ctypes.data), zb.WRAP_ZLIB, None)
            th.append(time.perf_counter() - t0)
            assert rc == 0 and not h_st.any() and bool((h_len == sz).all())
        okh = bool(np.array_equal(np.ctypeslib.as_array(C.cast(pin_dst, C.POINTER(C.c_uint8)), shape=(distinct * sz,)), host[:distinct * sz]))
        extra["inflate_batch_host_e2e"] = {"GBps": round(ns * sz / min(th) / 1e9, 3), "ms": round(min(th) * 1e3, 2), "host_buffers": "pinned",
                                           "h2d_bytes": zin, "d2h_bytes": ns * sz, "output_matches": okh}
        lib.dll.zb200_free_pinned(C.c_void_p(pin_z))
        # the e2e call again with malloc'ed (pageable) buffers -- what an unmodified caller of the reference passes
        pg_src = np.array(host, copy=True)
        pg_dst = np.empty(cap, dtype=np.uint8)
        pg_dst[:] = 0
        tp = []
        for _ in range(3):
            ol = C.c_ulong(cap)
            t0 = time.perf_counter()
            rc = lib.dll.compress2(C.c_void_p(pg_dst.ctypes.data), C.byref(ol), C.c_void_p(pg_src.ctypes.data), n, 1)
            tp.append(time.perf_counter() - t0)
            assert rc == 0 and ol.value == int(state["e2e_len"])
        extra["compress2_pageable_e2e"] = {"GBps": round(n / min(tp) / 1e9, 3), "ms": round(min(tp) * 1e3, 1), "host_buffers": "pageable (malloc)",
                                           "staging": "8 host threads, 4 MiB pieces through pinned slots"}
        del pg_src, pg_dst
        # uncompress() of ONE long stream: the compress2 output of the e2e leg (1 GiB in one zlib stream, pinned host
        # buffers); decoded in parallel at its chunk boundaries (BASELINE config 1's round trip, at config 2's size)
        step_e2e()                                             # pin_dst holds the level-1 stream again
        zlen1 = int(state["e2e_len"])
        pin_rt = lib.dll.zb200_alloc_pinned(n)
        assert pin_rt
        tu = []
        for _ in range(3):
            ul = C.c_ulong(n)
            t0 = time.perf_counter()
            rc = lib.dll.uncompress(C.c_void_p(pin_rt), C.byref(ul), C.c_void_p(pin_dst), zlen1)
            tu.append(time.perf_counter() - t0)
            assert rc == 0 and ul.value == n, (rc, lib.last_error())
        rt_ok = bool(np.array_equal(np.ctypeslib.as_array(C.cast(pin_rt, C.POINTER(C.c_uint8)), shape=(n,)), host))
        assert rt_ok, "uncompress(compress2(x)) != x"
        extra["uncompress_one_stream"] = {"GBps": round(n / min(tu) / 1e9, 3), "ms": round(min(tu) * 1e3, 1), "stream_bytes": zlen1,
                                          "host_buffers": "pinned", "round_trip_exact": rt_ok}
        lib.dll.zb200_free_pinned(C.c_void_p(pin_rt))
        # BASELINE config 1, direction (b): ONE 64 MiB text buffer compressed by the REFERENCE at level 6 (no flush points in
        # that stream: block starts are found), decoded by uncompress() of this library; the reference's own uncompress beside it
        if os.path.exists(zhelpers.REF_PATH):
            n1 = 64 << 20
            text = lib.synth(n1, kind=0, seed=7)
            refz = zhelpers.Ref()
            t0 = time.perf_counter()
            z1 = refz.compress2(text, 6)
            t_refc = time.perf_counter() - t0
            ob = C.create_string_buffer(n1)
            ul = C.c_ulong(n1)
            t0 = time.perf_counter()
            assert refz.dll.uncompress(ob, C.byref(ul), z1, len(z1)) == 0
            t_refu = time.perf_counter() - t0
            tz = []
            for _ in range(3):
                ul = C.c_ulong(n1)
                t0 = time.perf_counter()
                rc = lib.dll.uncompress(ob, C.byref(ul), z1, len(z1))
                tz.append(time.perf_counter() - t0)
                assert rc == 0 and ul.value == n1
            ok1 = ob.raw == text.tobytes()
            assert ok1, "uncompress(reference stream) differs"
            extra["uncompress_reference_stream_64MiB"] = {"GBps": round(n1 / min(tz) / 1e9, 3), "ms": round(min(tz) * 1e3, 1),
                                                          "stream_bytes": len(z1), "host_buffers": "pageable", "bit_exact": ok1,
                                                          "reference_uncompress_ms_one_core": round(t_refu * 1e3, 1),
                                                          "reference_compress2_level6_s_one_core": round(t_refc, 2)}
        # the full shape of BASELINE config 3: 100 000 streams of 64 KiB (the 2048 distinct ones repeated), 6.1 GiB out
        ns3 = 100000
        zs3 = (zs[:distinct] * ((ns3 + distinct - 1) // distinct))[:ns3]
        so3 = np.zeros(ns3 + 1, dtype=np.int64)
        so3[1:] = np.cumsum([len(z) for z in zs3])
        d_z3 = torch.from_numpy(np.frombuffer(b"".join(zs3) + b"\0" * 8, dtype=np.uint8).copy()).to(dev)
        d_so3 = torch.from_numpy(so3).to(dev)
        d_do3 = torch.arange(ns3 + 1, dtype=torch.int64, device=dev) * sz
        d_out3 = torch.empty(ns3 * sz, dtype=torch.uint8, device=dev)
        d_len3 = torch.zeros(ns3, dtype=torch.int64, device=dev)
        d_st3 = torch.zeros(ns3, dtype=torch.int32, device=dev)
        msi3, _ = timed(lambda: lib.inflate_batch_dev(d_z3.data_ptr(), d_so3.data_ptr(), ns3, d_out3.data_ptr(), d_do3.data_ptr(),
                                                      d_len3.data_ptr(), d_st3.data_ptr(), zb.WRAP_ZLIB, s), 3, 2)
        ok3 = int(d_st3.abs().sum()) == 0 and bool((d_len3 == sz).all()) and bool(torch.equal(d_out3[:distinct * sz], d_src[:distinct * sz]))
        extra["inflate_batch_config3"] = {"GBps": round(ns3 * sz / msi3 / 1e6, 3), "streams": ns3, "stream_bytes": sz, "ms": round(msi3, 2),
                                          "all_ok": ok3}
        del d_z3, d_out3
        # checksum at the per-GPU size of BASELINE config 4 scaled to one call's uInt limit: 4 GiB - 64 KiB
        n4 = (4 << 30) - 65536
        d_big = torch.empty(n4, dtype=torch.uint8, device=dev)
        d_big.view(torch.int64).random_()
        msc4, _ = timed(lambda: lib.checksum_dev(d_big.data_ptr(), n4, out2.data_ptr(), s), 5, 2)
        extra["crc32_adler32_fused_4GiB"] = {"GBps": round(n4 / msc4 / 1e6, 1), "ms": round(msc4, 4),
                                             "frac_of_hbm_peak": round(n4 / msc4 / 1e6 / peaks["hbm_gbs"], 4)}
        del d_big
        # ZIP archive (config 5 shape): files of log-uniform size 4 KiB .. 16 MiB cut from the same corpus, one
        # zb200_zip_build call on pinned host buffers (H2D + batch deflate + device-side assembly + one D2H)
        import random as _r
        rng = _r.Random(5)
        offs = [0]
        while offs[-1] < n and len(offs) <= 65535:
            offs.append(min(n, offs[-1] + int(4096 * 2 ** rng.uniform(0, 12))))
        nf = len(offs) - 1
        zoff = np.array(offs, dtype=np.uint64)
        znames = (C.c_char_p * nf)(*[b"f%05d.bin" % i for i in range(nf)])
        zcap = lib.dll.zb200_zip_bound(znames, C.c_void_p(zoff.ctypes.data), nf)
        pin_zip = lib.dll.zb200_alloc_pinned(zcap)
        assert pin_zip
        zt = []
        for _ in range(3):
            ol = C.c_size_t(zcap)
            t0 = time.perf_counter()
            rc = lib.dll.zb200_zip_build(znames, C.c_void_p(pin_src), C.c_void_p(zoff.ctypes.data), nf, 1, zb.DOS_DATETIME,
                                         C.c_void_p(pin_zip), C.byref(ol))
            zt.append(time.perf_counter() - t0)
            assert rc == 0, (rc, lib.last_error())
        zok = None
        if not args.no_verify:
            import io
            import zipfile
            arc = bytes(np.ctypeslib.as_array(C.cast(pin_zip, C.POINTER(C.c_uint8)), shape=(ol.value,)))
            zf = zipfile.ZipFile(io.BytesIO(arc))
            pick = rng.sample(range(nf), min(nf, 25))
            zok = all(zf.read("f%05d.bin" % i) == host[offs[i]:offs[i + 1]].tobytes() for i in pick) and len(zf.namelist()) == nf
            assert zok, "zip members do not read back"
        extra["zip_build_level1"] = {"GBps": round(n / min(zt) / 1e9, 3), "files": nf, "archive_bytes": int(ol.value),
                                     "ms": round(min(zt) * 1e3, 1), "host_buffers": "pinned", "members_read_back": zok}
        lib.dll.zb200_free_pinned(C.c_void_p(pin_zip))
        extra["reference_cpu"] = reference_side_paths()        # the reference's other paths on this box's host cores
        log(f"[rank 0] extras: {json.dumps(extra)}")

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": round(ms1, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"{args.size_mib} MiB synthetic mixed corpus per GPU (BASELINE.json configs[1]), deflate level 1, "
                                       "128 KiB chunks with 32 KiB dictionary priming, one zlib stream",
                           "level": 1, "chunk": 131072, "bytes_per_gpu": n, "l2": "inputs (1 GiB) exceed the 126 MB L2; no flush needed",
                           "parallelism": f"chunk-sharded x{world}" if world > 1 else "single GPU"},
                "ratio": round(n / clen1, 4), "compressed_bytes": int(clen1),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "verified_by_reference": verified, "extra": extra}
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
