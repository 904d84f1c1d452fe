"""One-shot deflate calls against the reference, byte for byte: random mixes of data kinds, lengths around the boundaries of
the CTA ranges of a long chunk's ordered phases (zb_deflate.cu: chain ranges, dfl_parse_multi_kernel), levels 4-9, every
strategy of the lazy path, window sizes and memLevels, preset dictionaries.  python tools/fuzz_onerun.py [seed] [trials]
— run under `timeout`."""
import ctypes as C
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

z = refz.ZlibBinding(zb.LIB_PATH, "")
ref = refz.ref()
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = random.Random(seed)
ZS = C.sizeof(refz.ZStream)


def whole(lib, d, level, wbits, mem, strat, dic):
    strm = refz.ZStream()
    assert lib.deflateInit2_(C.byref(strm), level, 8, wbits, mem, strat, lib.version, ZS) == 0
    if dic:
        assert lib.deflateSetDictionary(C.byref(strm), dic, len(dic)) == 0
    cap = lib.deflateBound(C.byref(strm), len(d)) + 64
    src, dst = C.create_string_buffer(d, max(len(d), 1)), C.create_string_buffer(cap)
    strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(d), C.addressof(dst), cap
    r = lib.deflate(C.byref(strm), refz.Z_FINISH)
    assert r == refz.Z_STREAM_END, r
    out = dst.raw[:cap - strm.avail_out]
    lib.deflateEnd(C.byref(strm))
    return out


def plain(n):
    parts, have = [], 0
    while have < n:
        k, m = rng.randrange(6), rng.choice((500, 30000, 262144, 900000))
        if k == 0:
            p = refz.gen(m, refz.GEN_TEXT, seed=rng.randrange(1 << 30))
        elif k == 1:
            p = refz.gen(m, refz.GEN_MARKOV, seed=rng.randrange(1 << 30))
        elif k == 2:
            p = rng.randbytes(m)
        elif k == 3:
            p = bytes(m)
        elif k == 4:
            p = (rng.randbytes(rng.randint(1, 300)) * (m // 100 + 1))[:m]
        else:
            p = refz.gen(m, refz.GEN_MIXED, seed=rng.randrange(1 << 30))
        parts.append(p); have += len(p)
    return b"".join(parts)[:n]


for t in range(trials):
    base = rng.choice((524288, 786432, 1048576, 2 * 1048576, 3 * 262144 * 3, 5 * 1048576))
    n = base + rng.choice((0, 0, 1, -1, 17, -4099, 131072, 262143))
    d = plain(n)
    level = rng.choice((4, 5, 6, 6, 7, 8, 9))
    strat = rng.choice((0, 0, 0, 1, 4))
    wb = rng.choice((15, 15, 15, -15, 31, 12, -10, 25))
    mem = rng.choice((8, 8, 8, 9, 1, 4))
    dic = refz.gen(rng.choice((100, 5000, 32768, 50000)), refz.GEN_MARKOV, seed=t) if rng.random() < 0.2 and wb < 16 else None
    got, want = whole(z, d, level, wb, mem, strat, dic), whole(ref, d, level, wb, mem, strat, dic)
    if got != want:
        k = next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), min(len(got), len(want)))
        print("MISMATCH trial %d: n %d level %d strategy %d windowBits %d memLevel %d dict %s: %d vs %d bytes, first difference at %d" %
              (t, n, level, strat, wb, mem, len(dic) if dic else None, len(got), len(want), k), flush=True)
        sys.exit(1)
print("fuzz_onerun seed %d: %d one-shot streams are the reference's byte for byte" % (seed, trials), flush=True)
