"""inflate() fed in random slice patterns (tiny ... huge, varying per call) over random streams (chunk sizes, flush kinds,
wrappers, trailing bytes, damage, truncation): the bytes, the consumed length and the verdict are the reference's
one-shot inflate's.  Run under `timeout`."""
import ctypes as C
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

z = refz.ZlibBinding(os.path.join(ROOT, os.environ["ZB_LIB"]) if os.environ.get("ZB_LIB") else zb.LIB_PATH, "")
ref = refz.ref()
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
base = refz.gen(9000000, refz.GEN_MARKOV, seed=4)


def sliced(lib, data, wrap, cap, pattern):
    s = refz.ZStream()
    assert lib.inflateInit2_(C.byref(s), refz._wbits(wrap), lib.version, C.sizeof(refz.ZStream)) == 0
    src = C.create_string_buffer(bytes(data), max(len(data), 1))
    dst = C.create_string_buffer(cap)
    fed = produced = 0
    ret = 0
    k = 0
    stalls = 0
    while True:
        step = min(pattern[k % len(pattern)], len(data) - fed)
        k += 1
        s.next_in, s.avail_in = C.addressof(src) + fed, step
        s.next_out, s.avail_out = C.addressof(dst) + produced, cap - produced
        before = (s.avail_in, s.avail_out)
        ret = lib.inflate(C.byref(s), 0)
        fed += step - s.avail_in
        produced = cap - s.avail_out
        if ret != 0 and not (ret == refz.Z_BUF_ERROR and fed < len(data)):
            break
        stalls = stalls + 1 if (s.avail_in, s.avail_out) == before and step == 0 else 0
        if stalls > 2 or (fed >= len(data) and ret == refz.Z_BUF_ERROR):
            break
    msg = s.msg.decode() if s.msg else ""
    tin = s.total_in
    lib.inflateEnd(C.byref(s))
    return ret, msg, dst.raw[:produced], tin


n_ok = 0
for trial in range(40):
    n = rng.choice((100000, 900000, 4000000))
    d = base[rng.randrange(0, len(base) - n):][:n]
    wrap = rng.choice((refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP))
    chunk = rng.choice((0, 3000, 60000, 262144, 1000000))
    flushes = [rng.choice((refz.Z_SYNC_FLUSH, refz.Z_FULL_FLUSH, refz.Z_FULL_FLUSH)) for _ in range(rng.randint(1, 4))]
    s = bytearray(ref.deflate_stream(d, rng.choice((1, 6)), 0, wrap, chunk, chunk_flush=flushes))
    kind = rng.randrange(4)
    if kind == 1:
        s[rng.randrange(len(s))] ^= 1 << rng.randrange(8)
    elif kind == 2:
        s = s[:rng.randrange(1, len(s))]
    elif kind == 3:
        s += rng.randbytes(rng.randint(1, 50))
    pattern = [rng.choice((1, 100, 5000, 16384, 70000, 300000, 1000000, 5000000)) for _ in range(rng.randint(1, 5))]
    want = ref.inflate_all(bytes(s), wrap, cap=n + 64)
    got = sliced(z, s, wrap, n + 64, pattern)
    tag = (trial, n, wrap, chunk, flushes, kind, pattern)
    if want[0] == refz.Z_STREAM_END:
        assert got[0] == refz.Z_STREAM_END and got[2] == want[2] and got[3] == want[3], (tag, got[0], got[1], len(got[2]), got[3], want[3])
    elif want[0] == refz.Z_DATA_ERROR:
        assert got[0] == refz.Z_DATA_ERROR, (tag, got[0], got[1], want[1])
        assert got[1] == want[1] or wrap == refz.WRAP_RAW or kind == 1, (tag, got[1], want[1])
        k = min(len(got[2]), len(want[2]))
        assert got[2][:k] == want[2][:k], tag
    else:                                            # truncated input: Z_BUF_ERROR / Z_OK, a prefix of the data
        assert got[0] in (refz.Z_OK, refz.Z_BUF_ERROR), (tag, got[0], got[1])
        assert want[2].startswith(got[2]) or got[2].startswith(want[2]), tag
    n_ok += 1
print("fuzz ok: %d streams" % n_ok)
