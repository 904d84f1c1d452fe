"""CPU-side checks of the drop-in boundary: the library loads and exports every
symbol include/zb200.h declares; without a GPU every compute entry point fails
loudly (no CPU fallback); the pure-host combine functions match the golden
vectors."""
import ctypes as C
import os
import re

import pytest

import refz
import zlib_wasm_b200 as zb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    L = C.CDLL(zb.LIB_PATH, mode=C.RTLD_LOCAL)
    names = _declared("zb200.h")
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(zb.ABI) == names, "python ABI table and header drifted apart"


def test_combine_matches_golden(golden):
    L = zb.lib()
    for c in golden["checksums"]["combine"]:
        assert L.zb200_crc32_combine(0x12345678, 0x9abcdef0, c["len2"]) == c["crc"]
        assert L.zb200_crc32_combine_gen(c["len2"]) == c["gen"]
        assert L.zb200_crc32_combine_op(0x12345678, 0x9abcdef0, c["gen"]) == c["crc"]
        assert L.zb200_adler32_combine(0x00c8012d, 0x11e60398, c["len2"]) == c["adler"]
    assert L.zb200_adler32_combine(1, 1, -1) == 0xffffffff


def test_no_cpu_fallback():
    L = zb.lib()
    if L.zb200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert L.zb200_create(0, C.byref(h)) == zb.ERR_NO_DEVICE
    assert b"no CPU path" in L.zb200_last_error()
    with pytest.raises(zb.ZB200Error):
        zb.Context(0)


def test_inflate_messages_are_the_reference_literals():
    L = zb.lib()
    o = refz.oracle()
    for i in range(22):
        assert L.zb200_inflate_msg(i) == o.c_inflate_msg(i)


def test_zlib_surface_exported():
    """Every function include/zb200_zlib.h declares is exported by the library
    (the reference's zlib.h names and the src/wasm_module.c exports)."""
    src = open(os.path.join(ROOT, "include", "zb200_zlib.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = src[src.rindex("#define inflateInit2"):]          # the prototypes follow the last init macro
    names = sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", body)) - {"inflateInit2", "defined", "sizeof"})
    L = C.CDLL(zb.LIB_PATH, mode=C.RTLD_LOCAL)
    missing = [n for n in names if not hasattr(L, n)]
    assert len(names) >= 45 and not missing, missing


def test_crc_table_is_the_reference_table():
    """get_crc_table() (crc32.c:549): the 256-entry table equals the one the byte-wise CRC definition gives."""
    L = C.CDLL(zb.LIB_PATH, mode=C.RTLD_LOCAL)
    L.get_crc_table.restype = C.POINTER(C.c_uint32)
    t = L.get_crc_table()
    # table[n] = crc of the single byte n with zero pre/post conditioning
    def raw(n):
        c = n
        for _ in range(8):
            c = (0xedb88320 ^ (c >> 1)) if c & 1 else c >> 1
        return c
    assert [t[n] for n in range(256)] == [raw(n) for n in range(256)]
    assert t[1] == 0x77073096 and t[255] == 0x2d02ef8d          # crc32.h:5-58 first row / last entry


def test_z_stream_layout_matches_reference():
    assert C.sizeof(refz.ZStream) == 112          # zlib.h:90-110 on LP64
    z = refz.ZlibBinding(zb.LIB_PATH, "")
    s = refz.ZStream()
    assert z.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, z.version, 111) == refz.Z_VERSION_ERROR
    assert z.inflateInit2_(C.byref(s), 15, b"0.9", C.sizeof(refz.ZStream)) == refz.Z_VERSION_ERROR
    if zb.lib().zb200_device_count() == 0:        # no GPU: init must fail loudly, never fall back
        assert z.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, z.version, C.sizeof(refz.ZStream)) == refz.Z_STREAM_ERROR
        assert b"no CPU path" in s.msg


def test_only_the_c_surfaces_are_exported():
    """The library is linked with a version script (csrc/libzb200.map, the counterpart of the reference's zlib.map):
    nothing but zb200_*, zlib_* and the zlib.h / gz* names leaves it — no C++ runtime symbols."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", zb.LIB_PATH], capture_output=True, text=True).stdout
    names = [ln.split()[-1] for ln in out.splitlines() if ln.strip()]
    assert len(names) > 100
    stray = [n for n in names if n.startswith("_Z") or n.startswith("__") or "std" in n]
    assert not stray, stray[:5]
