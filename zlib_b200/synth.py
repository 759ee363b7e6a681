"""Synthetic corpora of SURVEY.md section 8(d) (zlib_b200/libzbsynth.so, built from zlib_b200/synth/zbsynth.c).

kind 0 = text (T), 1 = mixed (M), 2 = xorshift64* noise.  Deterministic in (kind, seed, byte offset), generated page
by page (64 KiB) on the host cores.  Workload generator for benchmarks and tests -- deliberately a library of its own:
bench.py's reference arm uses it and must never map libzb200.so.
"""
import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libzbsynth.so")
PAGE = 65536
_dll = None


def _lib():
    global _dll
    if _dll is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        _dll = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
        _dll.zbsynth_fill.restype = C.c_int
        _dll.zbsynth_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64]
    return _dll


def fill(ptr: int, n: int, kind: int = 1, seed: int = 1, offset: int = 0) -> None:
    """Corpus bytes [offset, offset + n) into host memory at `ptr`; offset must be a multiple of 64 KiB."""
    if _lib().zbsynth_fill(C.c_void_p(ptr), n, kind, seed, offset) != 0:
        raise ValueError("offset must be a multiple of 65536")


def synth(n: int, kind: int = 1, seed: int = 1, offset: int = 0):
    """numpy uint8 array holding corpus bytes [offset, offset + n); any offset."""
    import numpy as np
    lead = offset % PAGE
    a = np.empty(n + lead, dtype=np.uint8)
    fill(a.ctypes.data, n + lead, kind, seed, offset - lead)
    return a[lead:]
