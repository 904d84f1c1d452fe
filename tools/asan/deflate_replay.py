import sys, random, ctypes as C
sys.path.insert(0, '/root/repo/tests')
import refz
L = C.CDLL('/tmp/libdef_emul_asan.so')
L.emul_deflate_chunk_dict.restype = C.c_long
L.emul_deflate_chunk_dict.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
libc = C.CDLL(None)
libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]; libc.free.argtypes = [C.c_void_p]
rng = random.Random(9)
runs = 0
for kind in range(5):
    for n in (0, 1, 2, 3, 5, 259, 4097, 70001, 140000):
        d = refz.gen(max(n, 1), kind, seed=3 + kind)[:n]
        for level, strat in ((0, 0), (1, 0), (3, 0), (6, 0), (6, 3), (9, 0)):
            for skip in (0, min(n, 1000)):
                src = libc.malloc(max(n, 1)); C.memmove(src, d, n)
                cap = n + n // 8 + 1024
                out = libc.malloc(cap)
                st = (C.c_uint32 * 2)()
                r = L.emul_deflate_chunk_dict(src, n, skip, level, strat, 1, out, cap, st)
                assert r >= 0, (kind, n, level, strat, skip, r)
                libc.free(src); libc.free(out)
                runs += 1
print("asan deflate runs", runs)
