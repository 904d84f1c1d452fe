import sys, random, ctypes as C
sys.path.insert(0, '/root/repo/tests')
import refz
L = C.CDLL('/tmp/libinf_emul_asan.so')
u64 = C.c_uint64
L.emul_inflate_rounds.argtypes = [C.c_void_p, u64, C.c_void_p, u64, C.c_int, C.c_int] + [C.POINTER(u64)] * 2 + [C.POINTER(C.c_uint32)] * 2 + [C.POINTER(u64)] * 2 + [C.POINTER(C.c_int)]
L.emul_inflate.argtypes = [C.c_void_p, u64, C.c_void_p, u64, C.c_int, u64, u64] + [C.POINTER(u64)] * 2 + [C.POINTER(C.c_uint32)] * 2 + [C.POINTER(u64)] * 2 + [C.POINTER(C.c_int)]
o = refz.oracle()
rng = random.Random(5)
libc = C.CDLL(None)
libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]; libc.free.argtypes = [C.c_void_p]
n_runs = 0
for kind, level in ((refz.GEN_MARKOV, 6), (refz.GEN_MIXED, 1), (refz.GEN_TEXT, 9)):
    d = refz.gen(200000, kind, seed=70 + kind)
    for wrap in (0, 2):
        s = o.deflate_stream(d, level, 0, wrap, 70000)
        cuts = [len(s)] + [rng.randrange(1, len(s)) for _ in range(40)]
        for k in cuts:
            data = bytearray(s[:k])
            if rng.random() < 0.3 and k > 10:
                data[rng.randrange(k)] ^= 1 << rng.randrange(8)
            # exact-size heap buffers (4-byte aligned, the word readers need the enclosing words): ASAN sees any read beyond
            kk = (k + 3) & ~3
            src = libc.malloc(kk); C.memmove(src, bytes(data) + b"\0" * (kk - k), kk)
            cap = len(d) + 8
            dst = libc.malloc(cap)
            iu, ol, cb, co = u64(), u64(), u64(), u64()
            ck, isz, kd = C.c_uint32(), C.c_uint32(), C.c_int()
            L.emul_inflate(src, k, dst, cap, wrap, 0, 0, C.byref(iu), C.byref(ol), C.byref(ck), C.byref(isz), C.byref(cb), C.byref(co), C.byref(kd))
            for lanes in (-1, 1000):
                L.emul_inflate_rounds(src, k, dst, cap, wrap, lanes, C.byref(iu), C.byref(ol), C.byref(ck), C.byref(isz), C.byref(cb), C.byref(co), C.byref(kd))
            libc.free(src); libc.free(dst)
            n_runs += 3
print("asan runs", n_runs)
