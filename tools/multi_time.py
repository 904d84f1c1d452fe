"""One process, all GPUs of the box through zb200_multi_* (pinned host buffers, end to end):
python tools/multi_time.py [GiB]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

L = zb.lib()
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30)) // (1 << 20) * (1 << 20)
m = C.c_void_p()
assert L.zb200_multi_create(None, 0, C.byref(m)) == 0, zb.last_error()
g = L.zb200_multi_count(m)
S = 262144
h_in = L.zb200_host_alloc(n)
cap = L.zb200_deflate_bound(n, S, zb.FRAME_GZIP_MEMBERS) + 4096
h_out = L.zb200_host_alloc(cap)
h_back = L.zb200_host_alloc(n + 64)
piece = 256 << 20
for off in range(0, n, piece):
    k = min(piece, n - off)
    C.memmove(h_in + off, refz.gen(k, refz.GEN_MARKOV, 5, first_block=off // 65536), k)


def timed(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


crc, adler = C.c_uint32(0), C.c_uint32(0)
dt = timed(lambda: L.zb200_multi_checksum_host(m, C.c_void_p(h_in), n, 3, 0, 1, C.byref(crc), C.byref(adler)))
print("%d GPUs, %.1f GiB pinned: crc32+adler32 %.1f GB/s (crc %08x)" % (g, n / 2**30, n / dt / 1e9, crc.value), flush=True)
olen = C.c_size_t(cap)
for level in (1, 6):
    def run():
        olen.value = cap
        assert L.zb200_multi_deflate_host(m, C.c_void_p(h_in), n, S, level, 0, zb.FRAME_GZIP_MEMBERS, 1, C.c_void_p(h_out), C.byref(olen), None, None) == 0, zb.last_error()
    dt = timed(run)
    print("  deflate L%d (256 KiB gzip members): %.1f GB/s end to end, ratio %.3f" % (level, n / dt / 1e9, n / olen.value), flush=True)
# the members of that file, discovered by one GPU, inflated by all
ctx = zb.Context(0)
nmax = n // S + 2
tab = (zb.Member * nmax)()
blen, nm, st = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
assert L.zb200_gunzip_host(ctx.handle, C.c_void_p(h_out), olen.value, C.c_void_p(h_back), n + 64, C.byref(blen), C.byref(st), tab, nmax, C.byref(nm)) == 0
assert st.value == 0 and blen.value == n
res = (zb.MemberResult * nm.value)()
dt = timed(lambda: L.zb200_multi_inflate_host(m, C.c_void_p(h_out), C.c_void_p(h_back), tab, nm.value, zb.WRAP_GZIP, 1, res))
ok = all(r.status == 0 for r in res) and C.string_at(h_back, 1 << 20) == C.string_at(h_in, 1 << 20) and \
    C.string_at(h_back + n - (1 << 20), 1 << 20) == C.string_at(h_in + n - (1 << 20), 1 << 20)
print("  inflate of %d members: %.1f GB/s end to end, bit_exact_ends=%s" % (nm.value, n / dt / 1e9, ok), flush=True)
