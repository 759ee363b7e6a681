"""zlib_b200 -- Python-side handle on libzb200.so (tests, benchmarks, multi-GPU driver).

The product is the C-ABI shared library `zlib_b200/libzb200.so` (zlib.h + zb200.h
entry points over hand-written sm_100a kernels).  This package only loads it with
ctypes and offers thin helpers; there is no Python or CPU implementation of any
codec arithmetic here, and importing `zlib_b200.binding` raises if the library has
not been built.
"""
from .binding import Lib, load, LIB_PATH  # noqa: F401

__all__ = ["Lib", "load", "LIB_PATH"]
