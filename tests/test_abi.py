"""CPU: the C-ABI library loads and exports every symbol that include/*.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, stop_marker=None):
    text = open(os.path.join(ROOT, "include", header)).read()
    if stop_marker and stop_marker in text:
        text = text[:text.index(stop_marker)]
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
        names.add(m.group(1))
    return names


def test_zlib_h_symbols_exported(lib):
    names = {n for n in _declared("zlib.h", "---- outside the hot path")
             if re.match(r"^(deflate|inflate|compress|uncompress|crc32|adler32|zlib|zError|get_crc)", n)}
    names -= {"deflateInit", "inflateInit", "deflateInit2", "inflateInit2"}      # macros
    assert len(names) >= 35
    for n in sorted(names):
        assert hasattr(lib.dll, n), f"{n} declared in include/zlib.h but not exported"


def test_zb200_h_symbols_exported(lib):
    names = {n for n in _declared("zb200.h") if n.startswith("zb200_")}
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(lib.dll, n), f"{n} declared in include/zb200.h but not exported"
    assert not lib.missing


def test_z_stream_layout():
    from zlib_b200.binding import z_stream
    assert C.sizeof(z_stream) == 112                      # h/zlib.h:82-101 on LP64
    offs = [getattr(z_stream, f).offset for f, _ in z_stream._fields_]
    assert offs == [0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104]


def test_header_compiles_as_c(tmp_path):
    """include/zlib.h is self-contained C89-compatible and agrees with the ctypes layout."""
    src = tmp_path / "t.c"
    src.write_text('#include "zlib.h"\n#include "zb200.h"\n#include <stddef.h>\n'
                   "int a[sizeof(z_stream)==112?1:-1]; int b[offsetof(z_stream,adler)==96?1:-1];\n"
                   "int main(void){z_stream s; s.zalloc=Z_NULL; (void)s; return deflateInit2(&s,6,Z_DEFLATED,15,8,0)==99;}\n")
    import subprocess
    subprocess.check_call(["gcc", "-std=c89", "-pedantic", "-Wall", "-Werror", "-Wno-long-long", "-c",
                           "-I", os.path.join(ROOT, "include"), str(src), "-o", str(tmp_path / "t.o")])


def test_misc_values(lib):
    assert lib.dll.zlibVersion() == b"1.2.3"
    assert lib.dll.zlibCompileFlags() == 0xA9          # reference build on LP64 (golden kat.json)
    assert lib.dll.compressBound(64 << 20) == 67129355
    assert lib.dll.compressBound(1 << 30) == 1074069515
    assert lib.dll.zError(-3) == b"data error"
    import json
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))
    for c1, c2, n, want in k["crc32_combine"]:
        assert lib.crc32_combine(c1, c2, n) == want
    for a1, a2, n, want in k["adler32_combine"]:
        assert lib.adler32_combine(a1, a2, n) == want
    assert lib.dll.crc32(0, None, 0) == 0 and lib.dll.adler32(0, None, 0) == 1


def test_no_cpu_fallback_without_gpu(lib):
    """On a box without a GPU the engine refuses to work instead of computing on the host."""
    if lib.dll.zb200_device_count() > 0:
        pytest.skip("a GPU is present")
    assert lib.dll.zb200_init(-1) != 0
    assert b"no CPU path" in lib.dll.zb200_last_error()
    rc, _ = lib.compress2(b"hello, hello!", 6)
    assert rc == -2                                       # Z_STREAM_ERROR, not a host-side result


def test_only_the_c_abi_is_exported():
    """Symbol hygiene (SURVEY.md hard part 9): the dynamic symbol table holds zlib.h, zb200.h and the three zutil
    names the reference's own callers reach for -- no C++ runtime instantiations, no mangled names."""
    import subprocess
    from zlib_b200 import LIB_PATH
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if l.strip()]
    assert names and not [n for n in names if n.startswith("_Z")], [n for n in names if n.startswith("_Z")][:5]
    allowed = ("zb200_", "deflate", "inflate", "compress", "uncompress", "crc32", "adler32", "get_crc_table",
               "zlib", "zError", "z_errmsg", "zcalloc", "zcfree")
    stray = [n for n in names if not n.startswith(allowed)]
    assert not stray, stray


def test_corpus_generator_is_its_own_library():
    """bench.py's reference arm generates its input without mapping libzb200.so (zlib_b200/libzbsynth.so)."""
    import numpy as np
    from zlib_b200 import synth
    a = synth.synth(300000, 1, 1)
    b = synth.synth(100000, 1, 1, offset=65536 + 11)
    assert np.array_equal(a[65536 + 11:65536 + 11 + 100000], b)
    assert not hasattr(__import__("zlib_b200").load().dll, "zb200_synth")
