"""GPU parity: one-warp-per-member inflate through the C ABI against golden
vectors, the reference's own streams, and the reference's error classes."""
import base64
import random

import pytest

import refz
import zlib_wasm_b200 as zb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zb.Context(0)
    yield c
    c.close()


def msg(st):
    return zb.lib().zb200_inflate_msg(st).decode()


def one(ctx, stream, wrap, cap):
    out, res = ctx.inflate_host(stream, [(0, len(stream), 0, cap)], wrap, 1, out_size=max(cap, 1))
    r = res[0]
    return r.status, out[:r.out_len], r


def test_puff_vectors(ctx, golden):
    for v in golden["puff_vectors"]:
        st, out, r = one(ctx, bytes.fromhex(v["hex"]), zb.WRAP_RAW, 4096)
        if v["ret"] == refz.Z_STREAM_END:
            assert st == 0 and out.hex() == v["out_hex"] and r.in_used == v["total_in"], v
        elif v["ret"] == refz.Z_BUF_ERROR:
            assert msg(st) == "truncated input", (v, msg(st))
        else:
            assert msg(st) == v["msg"], (v, msg(st))


def test_zeros_raw(ctx, golden):
    z = golden["zeros_raw"]
    st, out, r = one(ctx, base64.b64decode(z["b64"]), zb.WRAP_RAW, z["out_len"] + 7)
    assert st == 0 and len(out) == 1234567 and set(out) == {0} and r.check == z["crc32"] and r.in_used == z["total_in"]


def test_golden_streams_batched(ctx, golden):
    """All golden streams with inline bytes as ONE batch per wrapper kind."""
    for wrap in (0, 1, 2):
        es = [e for e in golden["streams"] if "hex" in e and e["wrap"] == wrap]
        blob, members, ooff = b"", [], 0
        for e in es:
            s = bytes.fromhex(e["hex"])
            members.append((len(blob), len(s), ooff, e["n"]))
            blob += s + b"\xaa" * (len(blob) % 3)      # odd gaps: members start at arbitrary alignment
            ooff += e["n"] + 5
        out, res = ctx.inflate_host(blob, members, wrap, 1, out_size=ooff + 8)
        for e, m, r in zip(es, members, res):
            assert r.status == 0, (e["input"], e["level"], msg(r.status))
            assert out[m[2]:m[2] + r.out_len] == refz.gen(e["n"], e["kind"]) and r.in_used == m[1]


@pytest.mark.parametrize("wrap", [0, 1, 2])
def test_reference_streams_multi_member(ctx, wrap):
    if not refz.have_ref():
        pytest.skip("compiled reference not available")
    ref = refz.ref()
    rng = random.Random(wrap)
    blob, members, datas, ooff = b"", [], [], 0
    for i in range(48):
        kind = rng.choice([refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_MIXED, refz.GEN_RANDOM, refz.GEN_BYTES])
        n = rng.choice([0, 1, 100, 65536, rng.randrange(65536, 1 << 20)])
        d = refz.gen(n, kind, seed=1000 + i)
        level, strategy = rng.choice([(1, 0), (6, 0), (9, 0), (6, 1), (6, 2), (1, 3), (6, 4), (4, 0)])
        s = ref.deflate_stream(d, level, strategy, wrap, rng.choice([0, 0, 70000]))
        members.append((len(blob), len(s), ooff, n))
        datas.append(d)
        blob += s
        ooff += n
    out, res = ctx.inflate_host(blob, members, wrap, 1, out_size=ooff + 1)
    o = refz.oracle()
    for m, d, r in zip(members, datas, res):
        assert r.status == 0, msg(r.status)
        assert r.out_len == len(d) and out[m[2]:m[2] + len(d)] == d and r.in_used == m[1]
        assert r.check == (o.adler32(d) if wrap == 1 else o.crc32(d))


def test_error_classes_match_reference(ctx):
    if not refz.have_ref():
        pytest.skip("compiled reference not available")
    ref = refz.ref()
    rng = random.Random(21)
    d = refz.gen(20000, refz.GEN_MIXED, seed=5)
    for wrap in (0, 1, 2):
        s = ref.deflate_stream(d, 6, 0, wrap, 0)
        cases, blob, members = [], b"", []
        for trial in range(150):
            b = bytearray(s)
            if trial % 3 == 0:
                b = b[:rng.randrange(0, len(b))]
            else:
                i = rng.randrange(0, len(b))
                b[i] ^= 1 << rng.randrange(8)
            cases.append(bytes(b))
            members.append((len(blob), len(b), trial * (len(d) + 64), len(d) + 64))
            blob += bytes(b)
        out, res = ctx.inflate_host(blob, members, wrap, 1, out_size=150 * (len(d) + 64))
        for b, m, r in zip(cases, members, res):
            ret, rmsg, rout, rin = ref.inflate_all(b, wrap, cap=len(d) + 64)
            if ret == refz.Z_STREAM_END:
                assert r.status == 0 and out[m[2]:m[2] + r.out_len] == rout
            elif ret == refz.Z_DATA_ERROR:
                assert msg(r.status) == rmsg, (msg(r.status), rmsg)
            else:
                assert msg(r.status) in ("truncated input", "output buffer full"), (ret, msg(r.status))


def test_output_full_and_resume(ctx):
    """A member cut short reports the last block boundary; feeding the rest and
    resuming there completes it (the mechanism behind the streaming inflate())."""
    o = refz.oracle()
    d = refz.gen(600000, refz.GEN_MARKOV, seed=77)
    s = o.deflate_stream(d, 6, 0, refz.WRAP_GZIP, 0)
    cut = len(s) // 2
    import ctypes as C
    L = zb.lib()
    out = C.create_string_buffer(len(d) + 16)
    m = (zb.Member * 1)(zb.Member(0, cut, 0, len(d) + 16, 0, 0))
    res = (zb.MemberResult * 1)()
    assert L.zb200_inflate_host(ctx.handle, s, out, m, 1, zb.WRAP_GZIP, 1, res) == 0, zb.last_error()
    assert msg(res[0].status) == "truncated input" and 0 < res[0].resume_out <= res[0].out_len
    assert out.raw[:res[0].out_len] == d[:res[0].out_len]
    m[0] = zb.Member(0, len(s), 0, len(d) + 16, res[0].resume_bit, res[0].resume_out)
    assert L.zb200_inflate_host(ctx.handle, s, out, m, 1, zb.WRAP_GZIP, 1, res) == 0, zb.last_error()
    assert res[0].status == 0 and res[0].out_len == len(d) and out.raw[:len(d)] == d and res[0].in_used == len(s)
    # output too small
    st, _, r = one(ctx, s, zb.WRAP_GZIP, len(d) - 1)
    assert msg(st) == "output buffer full"


def test_pipelined_host_entry_point(ctx):
    """Many members in pinned memory take the pipelined path of zb200_inflate_host: same
    bytes and per-member results as the plain path."""
    import ctypes as C
    import torch
    L = zb.lib()
    n, chunk = 2112 << 20, 1 << 20                      # > 1.5 GiB of output: two pieces
    d = refz.gen(n, refz.GEN_MARKOV, seed=8)
    d_in = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
    cap = L.zb200_deflate_bound(n, chunk, zb.FRAME_GZIP_MEMBERS)
    d_blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_end = torch.zeros(n // chunk, dtype=torch.int64, device="cuda")
    d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert L.zb200_deflate_dev(ctx.handle, C.c_void_p(d_in.data_ptr()), n, chunk, 6, 0, zb.FRAME_GZIP_MEMBERS, 1,
                               C.c_void_p(d_blob.data_ptr()), cap, C.c_void_p(d_end.data_ptr()), C.c_void_p(d_tot.data_ptr()), None) == 0
    torch.cuda.synchronize()
    ends, total = d_end.cpu().tolist(), int(d_tot.item())
    members, prev = [], 0
    for i, e in enumerate(ends):
        members.append(zb.Member(prev, e - prev, i * chunk, chunk, 0, 0))
        prev = e
    arr = (zb.Member * len(members))(*members)
    h_in, h_out = L.zb200_host_alloc(total), L.zb200_host_alloc(n)
    assert h_in and h_out
    try:
        C.memmove(h_in, d_blob[:total].cpu().numpy().tobytes(), total)
        res = (zb.MemberResult * len(members))()
        assert L.zb200_inflate_host(ctx.handle, C.c_void_p(h_in), C.c_void_p(h_out), arr, len(members), zb.WRAP_GZIP, 1, res) == 0, zb.last_error()
        assert all(r.status == 0 and r.out_len == chunk for r in res)
        assert C.string_at(h_out + (n - (64 << 20)), 64 << 20) == d[n - (64 << 20):] and C.string_at(h_out, 64 << 20) == d[:64 << 20]
        o = refz.oracle()
        assert res[7].check == o.crc32(d[7 * chunk:8 * chunk]) and res[7].in_used == members[7].in_len
    finally:
        L.zb200_host_free(C.c_void_p(h_in))
        L.zb200_host_free(C.c_void_p(h_out))
