"""inflate() through the zlib.h surface fed in slices of different sizes, this library next to the reference: a stream
WITHOUT flush points (the reference's one run: what compress2 / gzip / zpipe write) and this library's chunked form.
python tools/stream_time.py [MiB]"""
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

z = refz.ZlibBinding(zb.LIB_PATH, "")
ref = refz.ref() if refz.have_ref() else None
ctx = zb.Context(0)
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 128) << 20
d = refz.gen(n, refz.GEN_MARKOV, seed=3)
streams = {"one run (zlib.compress)": (zlib.compress(d, 6), refz.WRAP_ZLIB),
           "256 KiB full-flush chunks": (ctx.deflate_host(d, 6, 0, zb.FRAME_GZIP, 262144), refz.WRAP_GZIP)}
for what, (s, wrap) in streams.items():
    for sl in (16 << 10, 64 << 10, 256 << 10, 1 << 20, 16 << 20, None):
        row = []
        for name, lib in (("b200", z), ("reference", ref)):
            if lib is None:
                continue
            t0 = time.perf_counter()
            ret, m, out, tin = lib.inflate_all(s, wrap, cap=n + 64, in_slice=sl, out_slice=sl)
            dt = time.perf_counter() - t0
            row.append("%s %7.1f ms (%5.2f GB/s)%s" % (name, dt * 1e3, n / dt / 1e9, "" if ret == 1 and out == d else " WRONG"))
        print("%-26s slices of %-8s %s" % (what, "%d KiB" % (sl >> 10) if sl else "all", "   ".join(row)), flush=True)
