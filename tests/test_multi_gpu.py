"""GPU: the multi-GPU entry points of include/zb200.h driven by a plain C program (tests/c/multi_check.c, linked with
-lzb200 -lnccl): one zlib stream from all visible GPUs decoded bit-exact by the reference build, sharded checksums and
a sharded batch inflate checked against the reference.  On a one-GPU box the same code runs with one device."""
import os
import subprocess

import pytest

import zhelpers

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(zhelpers.ORACLE_DIR, "_ref")


def _checker():
    """The unmodified reference build when it is there, else the system zlib (an independent decoder either way)."""
    if os.path.exists(zhelpers.REF_PATH):
        return zhelpers.REF_PATH
    for p in ("/lib/x86_64-linux-gnu/libz.so.1", "/usr/lib/x86_64-linux-gnu/libz.so.1"):
        if os.path.exists(p):
            return p
    pytest.fail("no independent zlib to check against")


def test_c_program_uses_all_gpus(gpu_lib):
    exe = os.path.join(REFDIR, "multi_check")
    assert os.path.exists(exe), f"{exe} missing: run __graft_entry__.build()"
    ndev = gpu_lib.dll.zb200_multi_devices()
    assert ndev >= 1
    r = subprocess.run([exe, _checker(), str(ndev), "192"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "multi_check ok" in r.stdout
    assert f"on {ndev} GPU(s)" in r.stdout


def test_multi_entry_points_from_python(gpu_lib, oracle):
    """The same calls through ctypes with pageable buffers and odd sizes; gzip wrap; every device count up to what is there."""
    import ctypes as C
    import zlib
    data = gpu_lib.synth((40 << 20) + 777, kind=1, seed=21).tobytes()
    have = gpu_lib.dll.zb200_multi_devices()
    for ndev in sorted({1, have}):
        for wrap, wbits in ((1, 15), (2, 31), (0, -15)):
            cap = gpu_lib.compress_bound(len(data)) + 64
            out = C.create_string_buffer(cap)
            ol = C.c_size_t(cap)
            crc, adl = C.c_uint32(0), C.c_uint32(0)
            rc = gpu_lib.dll.zb200_multi_deflate(data, len(data), out, C.byref(ol), 1, wrap, ndev, C.byref(crc), C.byref(adl))
            assert rc == 0, gpu_lib.last_error()
            assert zlib.decompress(out.raw[:ol.value], wbits) == data
            assert crc.value == oracle.crc32(data) and adl.value == oracle.adler32(data)
        crc, adl = C.c_uint32(0), C.c_uint32(0)
        assert gpu_lib.dll.zb200_multi_checksum(data, len(data), ndev, C.byref(crc), C.byref(adl)) == 0
        assert crc.value == oracle.crc32(data) and adl.value == oracle.adler32(data)
