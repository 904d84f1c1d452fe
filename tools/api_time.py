"""compress2 / uncompress / crc32 through the zlib.h surface of libzb200.so next to the reference (config C1 and larger):
wall time of the calls themselves, pageable host buffers, after one warm-up call."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

z = refz.ZlibBinding(zb.LIB_PATH, "")
ref = refz.ref() if refz.have_ref() else None


def run(lib, d, level):
    n = len(d)
    cap = C.c_ulong(lib.compressBound(n))
    dst = C.create_string_buffer(cap.value)
    src = C.create_string_buffer(d, n)
    t0 = time.perf_counter()
    assert lib.compress2(dst, C.byref(cap), src, n, level) == 0
    t1 = time.perf_counter()
    back = C.create_string_buffer(n)
    bl = C.c_ulong(n)
    t2 = time.perf_counter()
    assert lib.uncompress(back, C.byref(bl), dst, cap.value) == 0 and bl.value == n
    t3 = time.perf_counter()
    c = lib.crc32_z(0, src, n)
    t4 = time.perf_counter()
    assert back.raw == d
    return t1 - t0, t3 - t2, t4 - t3, cap.value, c


for n in (1 << 20, 64 << 20, 512 << 20):
    d = refz.gen(n, refz.GEN_TEXT, seed=0x9E37)
    for name, lib in (("b200", z), ("reference", ref)):
        if lib is None:
            continue
        run(lib, d, 6)                                  # warm-up at the same size (device buffers grow once)
        tc, tu, tk, size, crc = run(lib, d, 6)
        print("%-9s %3d MiB  compress2(L6) %8.2f ms  uncompress %8.2f ms  crc32 %7.2f ms  -> %d bytes, crc %08x" %
              (name, n >> 20, tc * 1e3, tu * 1e3, tk * 1e3, size, crc), flush=True)
