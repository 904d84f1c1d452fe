"""Carried history (zb200.h ZB200_CHUNK_CARRY) against the reference, byte for byte: per chunk the reference's
deflateInit2(raw) + deflateSetDictionary(the 32 KiB before the chunk) + deflate(chunk, Z_SYNC_FLUSH | Z_FINISH), laid end to end
(deflate.c:550-632,1211-1218) — random mixes of data kinds, chunk sizes from far below the window to beyond the one-run
threshold, ragged tails, levels 4-9 and the literal-only / run-length strategies, window sizes, memLevels, a preset dictionary
ahead of the first chunk, open ends.  Levels 1-3: the reference's inflate gives the input back.
python tools/fuzz_carry.py [seed] [trials]   — run under `timeout`."""
import ctypes as C
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

ref = refz.ref()
L = zb.lib()
ctx = zb.Context(0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = random.Random(seed)
ZS = C.sizeof(refz.ZStream)


def ref_chunk(piece, hist, level, wbits, mem, strat, flush):
    strm = refz.ZStream()
    assert ref.deflateInit2_(C.byref(strm), level, 8, -wbits, mem, strat, ref.version, ZS) == 0
    if hist:
        assert ref.deflateSetDictionary(C.byref(strm), hist, len(hist)) == 0
    cap = len(piece) + len(piece) // 8 + 1024 + 10 * (len(piece) // 100)
    src, dst = C.create_string_buffer(piece, max(len(piece), 1)), C.create_string_buffer(cap)
    strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(piece), C.addressof(dst), cap
    r = ref.deflate(C.byref(strm), flush)
    assert (r == refz.Z_STREAM_END) if flush == refz.Z_FINISH else (r == 0 and strm.avail_in == 0 and strm.avail_out > 0), (r, flush)
    out = dst.raw[:cap - strm.avail_out]
    ref.deflateEnd(C.byref(strm))
    return out


def reference(base, dl, chunk, level, wbits, mem, strat, finish):
    n = len(base) - dl
    nch = (n + chunk - 1) // chunk
    if nch == 0 and finish:
        nch = 1
    out = []
    for c in range(nch):
        pos = dl + c * chunk
        out.append(ref_chunk(base[pos:pos + chunk], base[max(0, pos - 32768):pos], level, wbits, mem, strat,
                             refz.Z_FINISH if (finish and c == nch - 1) else refz.Z_SYNC_FLUSH))
    return b"".join(out)


def ours(base, dl, chunk, level, wbits, mem, strat, finish):
    cap = L.zb200_deflate_bound(len(base), chunk, zb.FRAME_RAW) + 64
    out = C.create_string_buffer(cap)
    olen = C.c_size_t(cap)
    o = zb.DeflateOpts(level, strat, wbits, mem, dl, 0)
    r = L.zb200_deflate_host_opts(ctx.handle, base, len(base), chunk, C.byref(o), zb.FRAME_RAW | zb.CHUNK_CARRY, finish, out, C.byref(olen), None, None, None)
    assert r == 0, zb.last_error()
    return out.raw[:olen.value]


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


exact = 0
for t in range(trials):
    n = rng.choice((0, 1, 300, 70000, 262144, 600000, 1500000, 4000000)) + rng.choice((0, 0, 1, 17, 4099))
    chunk = rng.choice((700, 4096, 20000, 32768, 32769, 65536, 100000, 262144, 524288, 700000, 1 << 20))
    if n // chunk > 3000:
        chunk = n // 3000 + 1
    level = rng.choice((1, 2, 3, 4, 5, 6, 6, 7, 8, 9))
    strat = rng.choice((0, 0, 0, 1, 2, 3, 4))
    wb = rng.choice((15, 15, 15, 14, 12, 10, 9))
    mem = rng.choice((8, 8, 8, 9, 1, 4))
    dl = rng.choice((0, 0, 0, 3, 5000, 1 << (wb - 1), 1 << wb))
    finish = 0 if rng.random() < 0.15 else 1
    base = plain(dl + n)
    got = ours(base, dl, chunk, level, wb, mem, strat, finish)
    what = "n %d chunk %d level %d strategy %d windowBits %d memLevel %d dict %d finish %d" % (n, chunk, level, strat, wb, mem, dl, finish)
    if level >= 4 or strat in (2, 3):
        want = reference(base, dl, chunk, level, wb, mem, strat, finish)
        if got != want:
            k = next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), min(len(got), len(want)))
            print("MISMATCH trial %d: %s: %d vs %d bytes, first difference at %d" % (t, what, len(got), len(want), k), flush=True)
            sys.exit(1)
        exact += 1
    stream = got + (b"" if finish else b"\x03\x00")            # (an open end: close it with an empty fixed block)
    err, msg, back, used = ref.inflate_all(stream, refz.WRAP_RAW, cap=n + 16, dictionary=base[:dl] if dl else None)
    if back != base[dl:] or used != len(stream):
        print("DECODE MISMATCH trial %d: %s: %r" % (t, what, msg), flush=True)
        sys.exit(1)
print("fuzz_carry seed %d: %d streams decode with the reference's inflate, %d of them are the reference's bytes" % (seed, trials, exact), flush=True)
