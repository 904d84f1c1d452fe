"""ZB200_EXACT_FAST next to the default greedy path: python tools/exact_time.py [MiB] — levels 1-3 on word text through
zb200_deflate_host (pinned buffers): size and the library's kernel time of both forms, bytes compared with the reference's
where oracle/_ref is present (a bounded prefix)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n, CH = mib << 20, 262144
L = zb.lib()
L.zb200_host_alloc.restype = C.c_void_p
ctx = zb.Context(0)
d = refz.gen(n, refz.GEN_MARKOV, seed=3)
h_in = L.zb200_host_alloc(n)
C.memmove(h_in, d, n)
cap = L.zb200_deflate_bound(n, CH, zb.FRAME_RAW)
h_out = L.zb200_host_alloc(cap)
for level in (1, 2, 3):
    row = []
    for flag in (0, zb.EXACT_FAST):
        olen = C.c_size_t(cap)

        def call():
            olen.value = cap
            r = L.zb200_deflate_host(ctx.handle, C.c_void_p(h_in), n, CH, level, 0, zb.FRAME_RAW | flag, 1, C.c_void_p(h_out), C.byref(olen), None, None)
            assert r == 0, zb.last_error()

        call()
        t0 = time.perf_counter(); call(); dt = time.perf_counter() - t0
        ctx.profile(True); ctx.profile_read(); call(); k = ctx.profile_read(); ctx.profile(False)
        row.append((olen.value, dt * 1e3, sum(v[0] for v in k.values()), k.get("dfl_fast_exact_kernel", (0, 0))[0]))
    (s0, e0, k0, _), (s1, e1, k1, x1) = row
    print("level %d, %d MiB of text in %d-byte chunks: default %d B, %.1f ms end to end (kernels %.1f) | exact %d B (%.4f x), %.1f ms = %.2f GB/s (kernels %.1f, of which the walk %.1f)" % (
        level, mib, CH, s0, e0, k0, s1, s1 / s0, e1, n / e1 / 1e6, k1, x1), flush=True)
