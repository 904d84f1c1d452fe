"""smoke(): one small invocation of every hot-path leg on cuda:0, each checked
against the oracle (oracle/liboracle.so — the checker, never the product)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run():
    import zlib_wasm_b200 as zb
    import refz
    o = refz.oracle()
    ctx = zb.Context(0)
    data = refz.gen(3 << 20, refz.GEN_MIXED)

    # checksums (both kernel shapes: < 8 MiB goes through ck_seg, >= 8 MiB through ck_big)
    for blob in (data, refz.gen(9 << 20, refz.GEN_BYTES)):
        crc, adler = ctx.checksum_host(blob)
        assert (crc, adler) == (o.crc32(blob), o.adler32(blob)), "checksum parity"

    # inflate: three gzip members compressed by the oracle's deflate restatement
    parts = [data[:700000], data[700000:1500000], data[1500000:]]
    blob, members, off, ooff = b"", [], 0, 0
    for p in parts:
        s = o.deflate_stream(p, 6, 0, refz.WRAP_GZIP, 0)
        members.append((off, len(s), ooff, len(p)))
        blob += s
        off += len(s)
        ooff += len(p)
    out, res = ctx.inflate_host(blob, members, zb.WRAP_GZIP, 1, out_size=len(data))
    assert all(r.status == 0 for r in res), [zb.lib().zb200_inflate_msg(r.status) for r in res]
    assert out[:len(data)] == data, "inflate parity"

    # deflate: every stream must decode (by the oracle's inflate) to the input
    if hasattr(ctx, "deflate_host"):
        for level in (1, 6):
            s = ctx.deflate_host(data, level, 0, zb.FRAME_ZLIB, 262144)
            err, msg, back, used = o.inflate_all(s, refz.WRAP_ZLIB, cap=len(data) + 8)
            assert err == 0 and back == data and used == len(s), ("deflate parity", level, msg)
            if level >= 4:   # levels 4..9 reproduce the reference byte for byte on the same chunking
                assert s == o.deflate_stream(data, level, 0, refz.WRAP_ZLIB, 262144), "deflate byte parity"
    # ONE run of blocks (what compress2 emits) by several CTAs: the oracle's own one-shot stream byte for byte; and that
    # stream — one member without flush points — decoded chunk by chunk at its block headers
    import ctypes as C
    one = ctx.deflate_host(data, 6, 0, zb.FRAME_ZLIB, len(data))
    assert one == o.deflate_stream(data, 6, 0, refz.WRAP_ZLIB, 0), "one-run byte parity"
    back = C.create_string_buffer(len(data) + 16)
    res1 = zb.MemberResult()
    r = zb.lib().zb200_inflate_stream_host(ctx.handle, one, len(one), zb.WRAP_ZLIB, back, len(data) + 16, C.byref(res1))
    assert r == 0 and res1.status == 0 and back.raw[:res1.out_len] == data, "chunk-parallel inflate parity"
    print("smoke ok: launches =", zb.lib().zb200_launch_count())
    ctx.close()
