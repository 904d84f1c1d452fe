"""Garbage tolerance of the block-parallel decode (csrc/zb_inflate_blocks.cuh) behind zb200_inflate_stream_host: streams
without flush points — intact, damaged, truncated, with garbage behind them, with their own block headers copied to
other places (true-looking false candidates), plain noise (stored-block look-alikes) — must come back (no hang, no
crash) with the status and bytes the one-member path gives for the same input, which the tests pin to the reference.
python tools/fuzz_blocks.py [seed] [trials]   — run under `timeout`."""
import ctypes as C
import os
import random
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

L = zb.lib()
ctx = zb.Context(0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rng = random.Random(seed)


def stream(blob, wrap, cap):
    out = C.create_string_buffer(max(cap, 1))
    res = zb.MemberResult()
    r = L.zb200_inflate_stream_host(ctx.handle, bytes(blob), len(blob), wrap, out, cap, C.byref(res))
    return r, res.status, out.raw[:min(res.out_len, cap)], res.in_used


def one(blob, wrap, cap):
    out, res = ctx.inflate_host(bytes(blob), [(0, len(blob), 0, cap)], wrap, 1, out_size=max(cap, 1))
    return 0, res[0].status, out[:min(res[0].out_len, cap)], res[0].in_used


def plain(n):
    parts = []
    while sum(map(len, parts)) < n:
        k = rng.randrange(6)
        m = rng.choice((3000, 40000, 200000, 700000))
        if k == 0:
            parts.append(refz.gen(m, refz.GEN_TEXT, seed=rng.randrange(1 << 30)))
        elif k == 1:
            parts.append(refz.gen(m, refz.GEN_MARKOV, seed=rng.randrange(1 << 30)))
        elif k == 2:
            parts.append(rng.randbytes(m))
        elif k == 3:
            parts.append(bytes(m))
        elif k == 4:
            parts.append((rng.randbytes(rng.randint(1, 40)) * (m // 20 + 1))[:m])
        else:
            parts.append(refz.gen(m, refz.GEN_MIXED, seed=rng.randrange(1 << 30)))
    return b"".join(parts)[:n]


WB = {refz.WRAP_RAW: -15, refz.WRAP_ZLIB: 15, refz.WRAP_GZIP: 31}
base = []
for _ in range(8):
    d = plain(rng.choice((200000, 1500000, 4000000)))
    wrap = rng.choice(list(WB))
    co = zlib.compressobj(rng.choice((1, 6, 9)), zlib.DEFLATED, WB[wrap], 8, rng.choice((0, 0, 0, 1, 2, 3, 4)))
    base.append((co.compress(d) + co.flush(), d, wrap))

via_blocks = 0
for t in range(trials):
    s, d, wrap = base[rng.randrange(len(base))]
    b = bytearray(s)
    kind = rng.randrange(7)
    if kind == 1:
        for _ in range(rng.randint(1, 4)):
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
    elif kind == 2:
        b = b[:rng.randrange(1, len(b))]
    elif kind == 3:
        b += rng.randbytes(rng.randint(1, 5000))
    elif kind == 4:                                          # a stretch of the stream (headers included) copied elsewhere
        for _ in range(rng.randint(1, 3)):
            a, ln, to = rng.randrange(len(b)), rng.choice((40, 300, 5000)), rng.randrange(len(b))
            b[to:to + ln] = b[a:a + ln][:max(0, len(b) - to)]
    elif kind == 5:                                          # noise: stored-block and header look-alikes
        b = bytearray(rng.randbytes(rng.choice((5000, 70000, 900000))))
        for _ in range(rng.randint(0, 30)):
            at = rng.randrange(0, len(b) - 8)
            ln = rng.randrange(65536)
            b[at:at + 4] = bytes((ln & 255, ln >> 8, (ln & 255) ^ 255, (ln >> 8) ^ 255))
        if wrap == refz.WRAP_ZLIB:
            b[0:2] = b"\x78\x9c"
        elif wrap == refz.WRAP_GZIP:
            b[0:10] = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03"
    elif kind == 6:                                          # too little room
        pass
    cap = len(d) + 64 if kind != 6 else rng.randrange(0, len(d))
    ctx.profile(True); ctx.profile_read()
    got = stream(b, wrap, cap)
    prof = ctx.profile_read(); ctx.profile(False)
    via_blocks += any("count" in k for k in prof)
    want = one(b, wrap, cap)
    ok = got[0] == 0 and got[1] == want[1]
    if ok and got[1] == 0:
        ok = got[2] == want[2] and got[3] == want[3]
    elif ok and zb.lib().zb200_inflate_msg(got[1]).decode() == "output buffer full":
        ok = True                                            # (the batch paths report the size needed and copy nothing)
    elif ok:
        k = min(len(got[2]), len(want[2]))
        ok = got[2][:k] == want[2][:k]
    if not ok:
        open("/tmp/fuzz_blocks_fail_%d_%d.bin" % (seed, t), "wb").write(bytes(b))
        print("MISMATCH trial %d kind %d wrap %d len %d: stream (r %d, status %d, %d bytes, used %d) one-member (status %d, %d bytes, used %d)" %
              (t, kind, wrap, len(b), got[0], got[1], len(got[2]), got[3], want[1], len(want[2]), want[3]), flush=True)
        sys.exit(1)
print("fuzz_blocks seed %d: %d trials agree with the one-member path (%d went through the chunk kernels)" % (seed, trials, via_blocks), flush=True)
