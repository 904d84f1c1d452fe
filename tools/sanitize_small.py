"""A small end-to-end pass of every kernel on the hot path (checksums, deflate at a greedy
and a lazy level, gzip members, inflate) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Inputs are a few hundred KiB so the instrumented run stays in the minutes.
(On the pool this was developed on compute-sanitizer is closed; the script then serves as a
quick round-trip check of all legs.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

ctx = zb.Context(0)
ok = True
for kind in (refz.GEN_MARKOV, refz.GEN_MIXED):
    d = refz.gen(300000 + 12345, kind, seed=11)
    for level in (1, 3, 6):
        s = ctx.deflate_host(d, level, 0, zb.FRAME_GZIP, 131072)
        out, res = ctx.inflate_host(s, [(0, len(s), 0, len(d))], zb.WRAP_GZIP, 1, out_size=len(d) + 8)
        good = res[0].status == 0 and out[:res[0].out_len] == d
        print("kind %d level %d: %d -> %d bytes, round trip %s" % (kind, level, len(d), len(s), good), flush=True)
        ok &= good
    crc, adler = ctx.checksum_host(d)
    ok &= crc == refz.oracle().crc32(d) and adler == refz.oracle().adler32(d)
# round 2: ONE run of blocks shared by several CTAs (chain ranges, parse links, block-end scan), and ONE member without
# flush points decoded chunk by chunk (candidate scan / validate, count + list passes, source pointers, jumping, gather)
import ctypes as C  # noqa: E402
import zlib  # noqa: E402
d = refz.gen(1400000 + 321, refz.GEN_MIXED, seed=12)
for level in (4, 6):
    s = ctx.deflate_host(d, level, 0, zb.FRAME_ZLIB, len(d))          # chunk = the whole input: one run
    good = s == zlib.compress(d, level)
    out = C.create_string_buffer(len(d) + 16)
    res = zb.MemberResult()
    r = zb.lib().zb200_inflate_stream_host(ctx.handle, s, len(s), zb.WRAP_ZLIB, out, len(d) + 16, C.byref(res))
    good = good and r == 0 and res.status == 0 and out.raw[:res.out_len] == d
    print("one run, level %d: %d -> %d bytes, the reference's bytes and back: %s" % (level, len(d), len(s), good), flush=True)
    ok &= good
print("ALL OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
