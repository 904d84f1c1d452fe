"""ONE member without flush points through zb200_inflate_stream_host (the block-parallel decode of
csrc/zb_inflate_blocks.cuh): wall time of the call on pageable host buffers next to the reference's inflate on one
core, and the library's per-kernel times."""
import ctypes as C
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

if os.environ.get("ZB_LIB"):
    zb.LIB_PATH = os.path.join(ROOT, os.environ["ZB_LIB"])
ctx = zb.Context(0)
sizes = [int(a) << 20 for a in sys.argv[1:]] or [1 << 20, 16 << 20, 64 << 20, 256 << 20]
for n in sizes:
    for gen, name in ((refz.GEN_TEXT, "text"), (refz.GEN_MIXED, "mixed")):
        d = refz.gen(n, gen, seed=0x9E37)
        s = zlib.compress(d, 6)
        out = C.create_string_buffer(n + 16)
        res = zb.MemberResult()

        def call():
            r = zb.lib().zb200_inflate_stream_host(ctx.handle, s, len(s), zb.WRAP_ZLIB, out, n + 16, C.byref(res))
            assert r == 0 and res.status == 0 and res.out_len == n, (r, res.status, res.out_len)

        call()
        assert out.raw[:n] == d
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
        ctx.profile(True); call(); prof = ctx.profile_read(); ctx.profile(False)
        t0 = time.perf_counter(); back = zlib.decompress(s); tr = time.perf_counter() - t0
        assert back == d
        print("%4d MiB %-5s  b200 %8.2f ms (%6.2f GB/s)   zlib one core %8.2f ms   kernels: %s" %
              (n >> 20, name, min(ts) * 1e3, n / min(ts) / 1e9, tr * 1e3,
               ", ".join("%s %.2f ms x%d" % (k.replace("_kernel", ""), v[0], v[1]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]))), flush=True)
