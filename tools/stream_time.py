import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import refz, zlib_wasm_b200 as zb
z = refz.ZlibBinding(zb.LIB_PATH, "")
ctx = zb.Context(0)
n = 256 << 20
d = refz.gen(n, refz.GEN_MARKOV, seed=3)
s = ctx.deflate_host(d, 6, 0, zb.FRAME_GZIP, 262144)
for sl in (16 << 20, 4 << 20, 1 << 20):
    t0 = time.perf_counter()
    ret, m, out, tin = z.inflate_all(s, refz.WRAP_GZIP, cap=n + 64, in_slice=sl)
    dt = time.perf_counter() - t0
    print("inflate() in slices of %d MiB: %.3f s  %.2f GB/s  ok=%s" % (sl >> 20, dt, n / dt / 1e9, ret == 1 and out == d), flush=True)
