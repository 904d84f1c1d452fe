import ctypes as C, os, sys, time, zlib
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import refz, zlib_wasm_b200 as zb
ctx = zb.Context(0)
text = refz.gen(16 << 20, refz.GEN_TEXT, seed=5)
rnd = refz.gen(8 << 20, refz.GEN_RANDOM, seed=6)
zeros = bytes(8 << 20)
cnt = np.arange(2 << 20, dtype=np.uint32).tobytes()
runs = (b"ab" * 50000 + b"x" * 100000 + bytes(300000)) * 10
delta = (np.cumsum(np.random.default_rng(1).integers(-3, 4, 8 << 20)) & 255).astype(np.uint8).tobytes()
for name, d in (("text", text), ("text+random", text + rnd + text), ("text+zeros", text + zeros + text), ("text+counters", text + cnt + text),
                ("text+runs", text + runs + text), ("text+delta", text + delta + text)):
    n = len(d)
    s = zlib.compress(d, 6)
    out = C.create_string_buffer(n + 16)
    res = zb.MemberResult()
    def call():
        r = zb.lib().zb200_inflate_stream_host(ctx.handle, s, len(s), zb.WRAP_ZLIB, out, n + 16, C.byref(res))
        assert r == 0 and res.status == 0 and res.out_len == n, (r, res.status, res.out_len)
    call(); assert out.raw[:n] == d
    ctx.profile(True); call(); prof = ctx.profile_read(); ctx.profile(False)
    print("%-14s %3d MiB -> %8d: %s" % (name, n >> 20, len(s), ", ".join("%s %.2f" % (k.replace("_kernel", ""), v[0]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6])), flush=True)
