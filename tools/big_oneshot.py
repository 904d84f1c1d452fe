import ctypes as C, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import refz, zlib_wasm_b200 as zb
z = refz.ZlibBinding(zb.LIB_PATH, ""); ref = refz.ref()
for n in ((1 << 30), (1 << 30) - 12345):
    d = refz.gen(n, refz.GEN_MIXED, seed=91)
    cap = z.compressBound(n)
    a, al = C.create_string_buffer(cap), C.c_ulong(cap)
    t0 = time.perf_counter(); r = z.compress2(a, C.byref(al), d, n, 4); t1 = time.perf_counter()
    print("b200 compress2 L4 of %d bytes: rc %d, %.1f ms, %d bytes" % (n, r, (t1 - t0) * 1e3, al.value), flush=True)
    if n == (1 << 30):
        b, bl = C.create_string_buffer(cap), C.c_ulong(cap)
        t0 = time.perf_counter(); r2 = ref.compress2(b, C.byref(bl), d, n, 4); t1 = time.perf_counter()
        print("reference: rc %d, %.1f ms, %d bytes, identical %s" % (r2, (t1 - t0) * 1e3, bl.value, a.raw[:al.value] == b.raw[:bl.value]), flush=True)
    back, kl = C.create_string_buffer(n), C.c_ulong(n)
    t0 = time.perf_counter(); r3 = z.uncompress(back, C.byref(kl), a, al.value); t1 = time.perf_counter()
    print("b200 uncompress: rc %d, %.1f ms, ok %s" % (r3, (t1 - t0) * 1e3, back.raw == d), flush=True)
