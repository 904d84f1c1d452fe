"""Debug helper: the damaged raw stream of fuzz_inflate_slices.py seed 2 / trial 39 through inflate() at several slice
sizes, first differing output byte against the reference's one-shot inflate."""
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
base = refz.gen(9000000, refz.GEN_MARKOV, seed=4)
rng = random.Random(2)
for trial in range(40):
    n = rng.choice((100000, 900000, 4000000))
    d = base[rng.randrange(0, len(base) - n):][:n]
    wrap = rng.choice((refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP))
    chunk = rng.choice((0, 3000, 60000, 262144, 1000000))
    flushes = [rng.choice((refz.Z_SYNC_FLUSH, refz.Z_FULL_FLUSH, refz.Z_FULL_FLUSH)) for _ in range(rng.randint(1, 4))]
    lvl = rng.choice((1, 6))
    if trial < 39:
        # keep the generator in step without compressing: lengths matter only for kinds 1 and 2
        s = bytearray(ref.deflate_stream(d, lvl, 0, wrap, chunk, chunk_flush=flushes))
    else:
        s = bytearray(ref.deflate_stream(d, lvl, 0, wrap, chunk, chunk_flush=flushes))
    kind = rng.randrange(4)
    if kind == 1:
        s[rng.randrange(len(s))] ^= 1 << rng.randrange(8)
    elif kind == 2:
        s = s[:rng.randrange(1, len(s))]
    elif kind == 3:
        s += rng.randbytes(rng.randint(1, 50))
    pattern = [rng.choice((1, 100, 5000, 16384, 70000, 300000, 1000000, 5000000)) for _ in range(rng.randint(1, 5))]
s = bytes(s)
want = ref.inflate_all(s, wrap, cap=n + 64)[2]
print("stream", len(s), "wrap", wrap, "chunk", chunk, "ref out", len(want), flush=True)


def run_schedule(name, sched):
    st = refz.ZStream()
    assert z.inflateInit2_(C.byref(st), refz._wbits(wrap), z.version, C.sizeof(refz.ZStream)) == 0
    src = C.create_string_buffer(s, len(s))
    cap = n + 64
    dst = C.create_string_buffer(cap)
    fed = produced = 0
    ret = 0
    for step in sched:
        step = min(step, len(s) - fed)
        st.next_in, st.avail_in = C.addressof(src) + fed, step
        st.next_out, st.avail_out = C.addressof(dst) + produced, cap - produced
        ret = z.inflate(C.byref(st), 0)
        fed += step - st.avail_in
        produced = cap - st.avail_out
        if ret not in (0, refz.Z_BUF_ERROR):
            break
    z.inflateEnd(C.byref(st))
    out = dst.raw[:produced]
    diffs = [i for i in range(min(len(out), len(want))) if out[i] != want[i]]
    print("%-28s ret %d len %d ndiff %d %s" % (name, ret, len(out), len(diffs), (diffs[:3], diffs[-3:]) if diffs else ""), flush=True)


marks = [i for i in range(len(s) - 3) if s[i:i + 4] == b"\x00\x00\xff\xff"]
print("markers at", marks[:8], "damage at 292954", flush=True)
dmg = 292954
run_schedule("1-byte around the damage", [dmg - 1500] + [1] * 3000 + [len(s)])
m0 = marks[len(marks) // 2]
run_schedule("1-byte around a marker", [m0 - 600] + [1] * 1200 + [len(s)])
run_schedule("1-byte at start and end", [1] * 2500 + [len(s) - 5000] + [1] * 2600)
run_schedule("7-byte all", [7] * (len(s) // 7 + 2))
