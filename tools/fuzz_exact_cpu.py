"""CPU fuzz of the exact deflate_fast walk (zb_deflate.cuh fast_exact_chunk, host replay tests/emul/def_emul.cpp) against the
reference: random mixes of data kinds (text, runs, binary, noise, repeats with long matches), sizes around the window slides,
levels 1-3, strategies 0 / 1 / 4, windowBits 9-15, memLevel 1-9, Z_FINISH / Z_FULL_FLUSH ends.
python tools/fuzz_exact_cpu.py [seed] [trials]   (no GPU needed)"""
import ctypes as C
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import test_emul  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rng = random.Random(seed)
L = test_emul._build("def_emul")
L.emul_deflate_chunk_opts.restype = C.c_long
L.emul_deflate_chunk_opts.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_size_t, C.POINTER(C.c_uint32)]
L.emul_set_exact_fast(1)
ref = refz.ref()
ZS = C.sizeof(refz.ZStream)


def reference(d, level, strat, wbits, mem, final):
    strm = refz.ZStream()
    assert ref.deflateInit2_(C.byref(strm), level, 8, -wbits, mem, strat, ref.version, ZS) == 0
    cap = len(d) + len(d) // 8 + 1024 + 8 * (len(d) // 100)
    src, dst = C.create_string_buffer(d, max(len(d), 1)), C.create_string_buffer(cap)
    strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(d), C.addressof(dst), cap
    r = ref.deflate(C.byref(strm), refz.Z_FINISH if final else refz.Z_FULL_FLUSH)
    assert r == (1 if final else 0) and strm.avail_in == 0, r
    out = dst.raw[:cap - strm.avail_out]
    ref.deflateEnd(C.byref(strm))
    return out


def plain(n):
    parts, have = [], 0
    while have < n:
        k, m = rng.randrange(7), rng.choice((50, 500, 5000, 30000, 70000))
        if k == 0:
            p = refz.gen(m, refz.GEN_TEXT, seed=rng.randrange(1 << 30))
        elif k == 1:
            p = refz.gen(m, refz.GEN_MARKOV, seed=rng.randrange(1 << 30))
        elif k == 2:
            p = rng.randbytes(m)
        elif k == 3:
            p = bytes([rng.randrange(256)]) * m
        elif k == 4:
            p = (rng.randbytes(rng.randint(1, 300)) * (m // 100 + 1))[:m]
        elif k == 5:
            p = refz.gen(m, refz.GEN_MIXED, seed=rng.randrange(1 << 30))
        else:
            p = bytes(rng.choice(b"ab") for _ in range(m))
        parts.append(p); have += len(p)
    return b"".join(parts)[:n]


for t in range(trials):
    n = rng.choice((0, 1, 5, 300, 5000, 32506, 32768, 65274, 65536, 98304, 131072, 200000, 262144, 400000)) + rng.choice((0, 0, 1, 2, 3, 261, 262, 263))
    d = plain(n)
    level, strat = rng.choice((1, 2, 3)), rng.choice((0, 0, 1, 4))
    wbits, mem = rng.choice((15, 15, 14, 12, 10, 9)), rng.choice((8, 8, 9, 7, 4, 1))
    final = rng.random() < 0.6
    want = reference(d, level, strat, wbits, mem, final)
    cap = n + n // 8 + 1024 + 8 * (n // 100)
    out, st = C.create_string_buffer(cap), (C.c_uint32 * 2)()
    r = L.emul_deflate_chunk_opts(d, n, 0, level, strat, wbits, mem, 1 if final else 0, out, cap, st)
    if r < 0 or out.raw[:r] != want:
        print("MISMATCH trial %d: n %d level %d strategy %d windowBits %d memLevel %d final %d: %d vs %d" % (t, n, level, strat, wbits, mem, final, r, len(want)), flush=True)
        sys.exit(1)
print("fuzz_exact_cpu seed %d: %d chunks are the reference's deflate_fast bytes" % (seed, trials), flush=True)
