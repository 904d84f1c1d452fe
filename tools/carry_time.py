"""History carried from chunk to chunk (ZB200_CHUNK_CARRY) next to independent chunks: python tools/carry_time.py MiB [chunk bytes]
— for levels 1, 6 and 9 on the level's bench generator: compressed size and the library's kernel time of both forms through
zb200_deflate_host on pinned buffers (the carried form works on S + 32 KiB positions per chunk: 12.5 % more chain / parse work at
256 KiB chunks), and the carried stream decoded again by zb200_inflate_stream_host (one run of blocks with sync points)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402,F401
import bench_legs as BL  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
CH = int(sys.argv[2]) if len(sys.argv) > 2 else BL.CHUNK
n = mib << 20
L = zb.lib()
ctx = zb.Context(0)
cap = L.zb200_deflate_bound(n, CH, zb.FRAME_RAW)
h_out = BL.host_alloc(L, cap)
h_back = BL.host_alloc(L, n + 64)
for level, gen in ((1, "markov"), (6, "mixed"), (9, "mixed")):
    host = BL.host_alloc(L, n)
    BL.fill(host, n, gen, 0)
    row = []
    for flag in (0, zb.CHUNK_CARRY):
        olen = C.c_size_t(cap)

        def call():
            olen.value = cap
            r = L.zb200_deflate_host(ctx.handle, C.c_void_p(host), n, CH, level, 0, zb.FRAME_RAW | flag, 1, C.c_void_p(h_out), C.byref(olen), None, None)
            assert r == 0, zb.last_error()

        call()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
        ctx.profile(True); ctx.profile_read(); call(); k = ctx.profile_read(); ctx.profile(False)
        ksum = sum(v[0] for v in k.values())
        row.append((olen.value, min(ts) * 1e3, ksum))
        if flag:
            res = zb.MemberResult()
            t0 = time.perf_counter()
            r = L.zb200_inflate_stream_host(ctx.handle, C.c_void_p(h_out), olen.value, zb.WRAP_RAW, C.c_void_p(h_back), n + 64, C.byref(res))
            dt = time.perf_counter() - t0
            same = r == 0 and res.status == 0 and res.out_len == n and C.string_at(h_back, n) == C.string_at(host, n)
            print("  level %d: carried stream decoded by zb200_inflate_stream_host: %s, %.1f ms" % (level, "input back" if same else "MISMATCH r=%d status=%d" % (r, res.status), dt * 1e3), flush=True)
    (s0, e0, k0), (s1, e1, k1) = row
    print("level %d %s %d MiB, %d-byte chunks: independent %d B, %.1f ms end to end, kernels %.1f ms | carried %d B (%.4f x), %.1f ms, kernels %.1f ms (%.3f x)" % (
        level, gen, mib, CH, s0, e0, k0, s1, s1 / s0, e1, k1, k1 / k0), flush=True)
    L.zb200_host_free(C.c_void_p(host))
