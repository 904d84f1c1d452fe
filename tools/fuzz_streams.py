"""Garbage-tolerance check of the discovery paths (zb200_gunzip_host, zb200_inflate_stream_host): random bytes
salted with member headers / flush markers, damaged and truncated real files.  Every call must return (no hang,
no crash) with a status; valid prefixes must be the reference's bytes.  Run under `timeout`."""
import ctypes as C
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

L = zb.lib()
ctx = zb.Context(0)
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
ref = refz.ref() if refz.have_ref() else refz.oracle()


def gunzip(blob, cap):
    out = C.create_string_buffer(max(cap, 1))
    olen, nm, st = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
    r = L.zb200_gunzip_host(ctx.handle, bytes(blob), len(blob), out, cap, C.byref(olen), C.byref(st), None, 0, C.byref(nm))
    return r, st.value, out.raw[:min(olen.value, cap)], nm.value


def stream(blob, wrap, cap):
    out = C.create_string_buffer(max(cap, 1))
    res = zb.MemberResult()
    r = L.zb200_inflate_stream_host(ctx.handle, bytes(blob), len(blob), wrap, out, cap, C.byref(res))
    return r, res.status, out.raw[:min(res.out_len, cap)]


base = refz.gen(1500000, refz.GEN_MARKOV, seed=9)
files = []
for k in range(6):
    parts = [base[rng.randrange(0, 1000000):][:rng.choice((0, 100, 30000, 200000))] for _ in range(rng.randint(1, 12))]
    files.append((b"".join(ref.deflate_stream(p, rng.choice((1, 6)), 0, refz.WRAP_GZIP, rng.choice((0, 20000))) for p in parts), b"".join(parts)))
n_calls = 0
for trial in range(60):
    kind = trial % 4
    if kind == 0:                                    # pure noise with look-alike headers and markers
        b = bytearray(rng.randbytes(rng.choice((100, 5000, 300000))))
        for _ in range(rng.randint(0, 40)):
            at = rng.randrange(0, max(1, len(b) - 12))
            b[at:at + 10] = rng.choice((b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03", b"\x00\x00\xff\xff\x00\x00\xff\xff\x00\x00"))
        b[0:3] = b"\x1f\x8b\x08"; b[3] = 0
        want = None
    else:
        f, plain = files[rng.randrange(len(files))]
        b = bytearray(f)
        if kind == 1:                                # damaged
            for _ in range(rng.randint(1, 4)):
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        elif kind == 2:                              # truncated
            b = b[:rng.randrange(1, len(b))]
        else:                                        # garbage appended / inserted between members
            b += rng.randbytes(rng.randint(1, 3000))
        want = plain
    r, st, out, nm = gunzip(b, 4000000)
    assert r in (0, zb.ERR_OUTPUT, zb.ERR_PARAM), (trial, r, zb.last_error())
    if want is not None and r == 0:
        assert want.startswith(out) or kind == 1, (trial, kind, st, len(out))
    r, st, out = stream(b, refz.WRAP_GZIP, 4000000)
    assert r in (0, zb.ERR_OUTPUT, zb.ERR_PARAM), (trial, r)
    if want is not None and r == 0 and st == 0:
        assert want.startswith(out) or kind == 1
    n_calls += 2
print("fuzz ok: %d calls" % n_calls)
