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
                # Z_BUF_ERROR (or Z_OK with input exhausted): the reference ran out of input, or out of output room
                # (inflate.c:1259-1261); tell the two apart by which of them the reference had left
                assert ret in (refz.Z_BUF_ERROR, refz.Z_OK), ret
                want = "output buffer full" if len(rout) >= len(d) + 64 else "truncated input"
                assert msg(r.status) == want, (ret, msg(r.status), want, len(rout), rin, len(b))


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


def _complete_lengths(rng, n_used, max_len):
    """Code lengths of a random complete prefix code over n_used symbols (Kraft sum 1):
    leaves are split at random until there are enough of them."""
    leaves = [0]
    while len(leaves) < n_used:
        cand = [i for i, d in enumerate(leaves) if d < max_len]
        i = rng.choice(cand) if rng.random() < 0.5 else max(cand, key=lambda k: -leaves[k] + rng.random())
        d = leaves.pop(i)
        leaves += [d + 1, d + 1]
    rng.shuffle(leaves)
    return leaves


def _skewed_lengths(rng, n_used):
    """Huffman lengths of Zipf-like frequencies; None when a code comes out longer than 15 bits."""
    import heapq
    a = rng.uniform(0.4, 1.6)
    ranks = list(range(1, n_used + 1))
    rng.shuffle(ranks)
    f = sorted((1.0 / r ** a, i) for i, r in enumerate(ranks))
    heap = [(w, i, None) for w, i in f]
    heapq.heapify(heap)
    k = n_used
    while len(heap) > 1:
        a, b = heapq.heappop(heap), heapq.heappop(heap)
        heapq.heappush(heap, (a[0] + b[0], k, (a, b)))
        k += 1
    lens = [0] * n_used

    def walk(node, d):
        if node[2] is None:
            lens[node[1]] = max(d, 1)
        else:
            walk(node[2][0], d + 1); walk(node[2][1], d + 1)
    walk(heap[0], 0)
    return lens if max(lens) <= 15 else None


def test_warp_table_builder(ctx):
    """csrc/zb_inflate_tables.cuh (every table entry finds its symbol) against the serial
    construction that follows inftrees.c:32-299: same status, identical tables."""
    import ctypes as C
    rng = random.Random(11)
    cases = []

    def add(lit, dist):
        assert 257 <= len(lit) <= 286 and 1 <= len(dist) <= 30
        cases.append((list(lit), list(dist)))

    def spread(lens, n, must=None):
        out = [0] * n
        where = rng.sample(range(n), len(lens))
        if must is not None and must not in where:
            where[0] = must
        for w, l in zip(where, lens):
            out[w] = l
        return out

    for _ in range(1500):
        nlen, ndist = rng.randint(257, 286), rng.randint(1, 30)
        nu = rng.randint(2, nlen)
        lit = spread(_complete_lengths(rng, nu, 15), nlen, must=256)
        du = rng.randint(1, ndist)
        dist = spread(_complete_lengths(rng, du, 15), ndist) if du > 1 else spread([1], ndist)
        add(lit, dist)
    for _ in range(300):                                     # Huffman-shaped codes (the common case)
        nlen, ndist = rng.randint(257, 286), rng.randint(2, 30)
        lit = _skewed_lengths(rng, nlen) or spread(_complete_lengths(rng, nlen, 15), nlen)
        dist = _skewed_lengths(rng, ndist) or spread(_complete_lengths(rng, ndist, 15), ndist)
        add(lit, dist)
    for _ in range(300):                                     # damaged: one length changed -> over-subscribed or incomplete
        lit, dist = cases[rng.randrange(1500)]
        lit, dist = list(lit), list(dist)
        tgt = lit if rng.random() < 0.6 else dist
        tgt[rng.randrange(len(tgt))] = rng.randint(0, 15)
        add(lit, dist)
    add([15] * 257 + [0] * 29, [0] * 30)                    # incomplete, no distance codes
    add([0] * 256 + [1] + [0] * 29, [1])                    # the single one-bit codes inftrees.c:131-132 lets pass
    add([0] * 256 + [1, 1], [0, 1])
    add([0] * 256 + [2] + [0] * 29, [2])                    # single codes of two bits: rejected
    add([8] * 144 + [9] * 112 + [7] * 24 + [8] * 6, [5] * 30)   # the fixed code's shape (286 / 30 symbols)
    add([9] * 254 + [8] * 3 + [9] * 29, [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15])
    add([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15] + [0] * 240 + [0] + [0] * 10, [1])   # 256 unused: still a table
    n = len(cases)
    lens = bytearray(n * 320)
    counts = (C.c_uint32 * (2 * n))()
    for i, (lit, dist) in enumerate(cases):
        lens[i * 320:i * 320 + len(lit)] = bytes(lit)
        lens[i * 320 + len(lit):i * 320 + len(lit) + len(dist)] = bytes(dist)
        counts[2 * i], counts[2 * i + 1] = len(lit), len(dist)
    verdict = (C.c_uint32 * n)()
    buf = (C.c_uint8 * len(lens)).from_buffer(lens)
    r = zb.lib().zb200_selftest_tables(ctx.handle, C.addressof(buf), C.addressof(counts), n, C.addressof(verdict))
    assert r == 0, zb.last_error()
    bad = [(i, hex(v)) for i, v in enumerate(verdict) if v]
    assert not bad, (len(bad), bad[:10], [cases[i] for i, _ in bad[:2]])


def test_members_with_preset_dictionaries(ctx):
    """zb200_member.dict_len: the dictionary lies just before the member's output (inflate.c:1278-1312);
    raw and zlib members made by the reference with deflateSetDictionary, several per call, each with its own."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    base = refz.gen(900000, refz.GEN_TEXT, seed=31)
    for wrap in (refz.WRAP_RAW, refz.WRAP_ZLIB):
        blob, members, want, out_buf, opos = b"", [], [], bytearray(), 0
        for k, (dl, n, lvl) in enumerate(((1, 1000, 6), (777, 300000, 6), (32768, 200000, 9), (20000, 70000, 1), (32768, 0, 6))):
            dic = base[k * 50000:k * 50000 + dl]
            data = base[k * 50000 + dl // 2:k * 50000 + dl // 2 + n]
            s = ref.deflate_stream(data, lvl, 0, wrap, 0, dictionary=dic)
            out_buf += dic
            opos += dl
            members.append((len(blob), len(s), opos, n + 16, 0, 0, dl))
            want.append((opos, data))
            out_buf += bytes(n + 16)
            opos += n + 16
            blob += s
        out, res = ctx.inflate_host(blob, members, wrap, 1, out_size=opos, prefill=bytes(out_buf))
        for (off, data), r, m in zip(want, res, members):
            assert r.status == 0 and r.out_len == len(data) and r.in_used == m[1], (wrap, msg(r.status), r.out_len, len(data))
            assert out[off:off + len(data)] == data
        if wrap == refz.WRAP_ZLIB:                           # without the dictionary: the reference's Z_NEED_DICT, DICTID reported
            m = members[1]
            out, res = ctx.inflate_host(blob, [(m[0], m[1], 0, m[3])], wrap, 1)
            dictid = int.from_bytes(blob[m[0] + 2:m[0] + 6], "big")
            assert msg(res[0].status) == "need dictionary" and res[0].check == dictid and res[0].out_len == 0


def _gunzip(ctx, blob, cap, max_members=0):
    import ctypes as C
    L = zb.lib()
    out = C.create_string_buffer(max(cap, 1))
    olen, nm, st = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
    tab = (zb.Member * max(max_members, 1))()
    r = L.zb200_gunzip_host(ctx.handle, bytes(blob), len(blob), out, cap, C.byref(olen), C.byref(st),
                            tab if max_members else None, max_members, C.byref(nm))
    return r, st.value, out.raw[:min(olen.value, cap)], olen.value, nm.value, list(tab)[:min(nm.value, max_members)]


def test_gunzip_discovers_members(ctx):
    """zb200_gunzip_host: a gzip file's members found without an index (gzread.c:76-234 walks them serially),
    incl. look-alike headers inside stored data, optional header fields, and bytes after the last member."""
    ref = refz.ref() if refz.have_ref() else refz.oracle()
    rng = random.Random(17)
    base = refz.gen(3000000, refz.GEN_MARKOV, seed=5)
    noise = bytearray(refz.gen(400000, refz.GEN_RANDOM, seed=6))
    for at in range(1000, len(noise) - 100, 37000):            # look-alike member headers inside incompressible data
        noise[at:at + 10] = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03"
    pieces, blob = [], b""
    for k in range(300):
        n = rng.choice((0, 1, 77, 5000, 70000, 200000))
        d = base[rng.randrange(0, len(base) - n):][:n]
        s = ref.deflate_stream(d, rng.choice((1, 6, 9)), 0, refz.WRAP_GZIP, 0)
        if k % 50 == 7:                                         # FNAME + FCOMMENT + FEXTRA + FHCRC (RFC 1952 2.3)
            import zlib as pz
            hdr = bytearray(s[:10]); hdr[3] = 0x1e
            extra = b"\x04\x00ab\x01\x02" + b"name.txt\x00" + b"a comment\x00"
            h = bytes(hdr) + extra
            s = h + (pz.crc32(h) & 0xffff).to_bytes(2, "little") + s[10:]
        pieces.append((d, len(s)))
        blob += s
    for lvl in (0, 6):                                          # stored blocks carry the look-alikes verbatim
        s = ref.deflate_stream(bytes(noise), lvl, 0, refz.WRAP_GZIP, 0)
        pieces.append((bytes(noise), len(s)))
        blob += s
    want = b"".join(d for d, _ in pieces)
    r, st, out, need, nm, tab = _gunzip(ctx, blob, len(want) + 100, max_members=len(pieces) + 8)
    assert r == 0 and st == 0 and nm == len(pieces) and out == want, (r, msg(st), nm, len(pieces), need, len(want))
    off = opos = 0
    for (d, cl), m in zip(pieces, tab):
        assert (m.in_off, m.in_len, m.out_off, m.out_cap) == (off, cl, opos, len(d))
        off += cl; opos += len(d)
    # too small an output buffer: the size needed comes back
    r, st, out, need, nm, tab = _gunzip(ctx, blob, 1000)
    assert r == zb.ERR_OUTPUT and need == len(want)
    # bytes after the last member that are no member are ignored; a member cut short is reported
    r, st, out, need, nm, tab = _gunzip(ctx, blob + b"\x00" * 5000 + b"junk", len(want) + 100)
    assert r == 0 and st == 0 and out == want and nm == len(pieces)
    cut = len(blob) - pieces[-1][1] // 2
    r, st, out, need, nm, tab = _gunzip(ctx, blob[:cut], len(want) + 100)
    assert r == 0 and msg(st) == "truncated input" and nm == len(pieces) - 1 and out == want[:len(want) - len(noise)]
    bad = bytearray(blob); bad[pieces[0][1] + pieces[1][1] + 30] ^= 0x10
    r, st, out, need, nm, tab = _gunzip(ctx, bytes(bad), len(want) + 100)
    assert r == 0 and st != 0 and nm <= 2 and out == want[:len(out)]
    r, st, out, need, nm, tab = _gunzip(ctx, b"not a gzip file at all....", 100)
    assert r == 0 and msg(st) == "incorrect header check" and out == b""


def _stream(ctx, blob, wrap, cap):
    import ctypes as C
    out = C.create_string_buffer(max(cap, 1))
    res = zb.MemberResult()
    r = zb.lib().zb200_inflate_stream_host(ctx.handle, bytes(blob), len(blob), wrap, out, cap, C.byref(res))
    assert r == 0, zb.last_error()
    return res, out.raw[:min(res.out_len, cap)]


def test_stream_decoded_in_parallel_at_flush_points(ctx):
    """zb200_inflate_stream_host: full-flush runs in parallel, sync-flush runs merged (they reach behind their
    start), look-alike markers inside stored data merged, the reference's statuses for damaged input."""
    ref = refz.ref() if refz.have_ref() else refz.oracle()
    d = refz.gen(6000000, refz.GEN_MARKOV, seed=61)
    noise = bytearray(refz.gen(300000, refz.GEN_RANDOM, seed=62))
    for at in range(500, len(noise) - 8, 9973):
        noise[at:at + 4] = b"\x00\x00\xff\xff"
    mixed = d[:1000000] + bytes(noise) + d[1000000:2000000]
    for wrap in (refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP):
        for data, level, chunk in ((d, 6, 262144), (d, 1, 100000), (mixed, 6, 65536), (mixed, 0, 50000), (d[:70000], 6, 0)):
            s = ref.deflate_stream(data, level, 0, wrap, chunk)
            res, out = _stream(ctx, s, wrap, len(data) + 16)
            assert res.status == 0 and out == data and res.in_used == len(s), (wrap, level, chunk, msg(res.status), res.out_len, len(data))
            res, out = _stream(ctx, s + b"tail", wrap, len(data) + 16)
            assert res.status == 0 and res.in_used == len(s) and out == data
    if refz.have_ref():                                       # sync flushes: the runs need their predecessors
        s = refz.ref().deflate_stream(d[:3000000], 6, 0, refz.WRAP_ZLIB, 200000, chunk_flush=refz.Z_SYNC_FLUSH)
        res, out = _stream(ctx, s, refz.WRAP_ZLIB, 3000016)
        assert res.status == 0 and out == d[:3000000]
    s = ref.deflate_stream(d, 6, 0, refz.WRAP_GZIP, 262144)
    res, out = _stream(ctx, s, refz.WRAP_GZIP, 1000)          # too small: the size needed
    assert msg(res.status) == "output buffer full" and res.out_len == len(d)
    res, out = _stream(ctx, s[:len(s) // 2], refz.WRAP_GZIP, len(d) + 16)
    assert msg(res.status) == "truncated input" and d.startswith(out) and len(out) > len(d) // 3
    bad = bytearray(s); bad[len(s) // 3] ^= 0x40
    res, out = _stream(ctx, bytes(bad), refz.WRAP_GZIP, len(d) + 16)
    one, res1 = ctx.inflate_host(bytes(bad), [(0, len(bad), 0, len(d) + 16)], refz.WRAP_GZIP, 1)
    assert res.status == res1[0].status and res.status != 0
    bad = bytearray(s); bad[-6] ^= 1
    res, out = _stream(ctx, bytes(bad), refz.WRAP_GZIP, len(d) + 16)
    assert msg(res.status) == "incorrect data check"


def test_stream_runs_random_flush_mixtures(ctx):
    """Random chunk sizes and flush kinds (sync / partial / full mixed): whatever the chain of runs looks like,
    the stream decodes to the reference's bytes."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    rng = random.Random(23)
    for trial in range(24):
        kind = rng.choice((refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_MIXED))
        n = rng.choice((70000, 300000, 1200000))
        d = refz.gen(n, kind, seed=100 + trial)
        chunk = rng.choice((60, 700, 5000, 40000, 200000))
        if n // chunk > 4000:
            chunk = n // 4000
        flushes = [rng.choice((refz.Z_SYNC_FLUSH, refz.Z_FULL_FLUSH, refz.Z_FULL_FLUSH, refz.Z_PARTIAL_FLUSH)) for _ in range(rng.randint(1, 7))]
        wrap = rng.choice((refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP))
        level = rng.choice((0, 1, 6, 9))
        s = ref.deflate_stream(d, level, rng.choice((0, 0, 1, 2, 3, 4)), wrap, chunk, chunk_flush=flushes)
        res, out = _stream(ctx, s, wrap, n + 16)
        assert res.status == 0 and out == d and res.in_used == len(s), (trial, kind, n, chunk, flushes, wrap, level, msg(res.status), res.out_len)


def test_one_member_decoded_in_parallel_at_block_headers(ctx):
    """A member without flush points (what compress2 / gzip write): its dynamic block headers are found on the device,
    the chunks between them are counted, placed, decoded at once and the matches resolved by pointer jumping
    (csrc/zb_inflate_blocks.cuh).  The original bytes, the stream's length, the reference's verdict on damage."""
    import zlib
    cases = [(refz.GEN_TEXT, 1 << 20, 6, 15), (refz.GEN_MARKOV, 5000000, 6, 31), (refz.GEN_MIXED, 6000000, 9, -15),
             (refz.GEN_TEXT, 300000, 1, 15), (refz.GEN_MIXED, 3000000, 1, 31), (refz.GEN_MARKOV, 40 << 20, 6, 15),
             (refz.GEN_TEXT, 2500000, 4, -15)]
    for gen, n, level, wbits in cases:
        d = refz.gen(n, gen, seed=n ^ level)
        co = zlib.compressobj(level, zlib.DEFLATED, wbits)
        s = co.compress(d) + co.flush()
        wrap = refz.WRAP_RAW if wbits < 0 else refz.WRAP_GZIP if wbits > 15 else refz.WRAP_ZLIB
        ctx.profile(True)
        res, out = _stream(ctx, s, wrap, n + 16)
        prof = ctx.profile_read()
        ctx.profile(False)
        assert res.status == 0 and res.out_len == n and out == d and res.in_used == len(s), (gen, n, level, wbits, msg(res.status), res.out_len)
        assert any(k.startswith("inflate_") and "count" in k for k in prof), (sorted(prof), gen, n, level)   # the chunk path decoded it
        res, out = _stream(ctx, s + b"tail", wrap, n + 16)
        assert res.status == 0 and res.in_used == len(s) and out == d
    # adversarial: real block headers that are NOT block starts — a compressed stream carried as DATA inside stored blocks
    # (level 0), inside a literal-only stream (Z_HUFFMAN_ONLY) and inside a normal one: candidates the chain steps over
    inner = zlib.compress(refz.gen(2000000, refz.GEN_TEXT, seed=5), 6)
    payload = refz.gen(300000, refz.GEN_MARKOV, seed=6) + inner + refz.gen(700000, refz.GEN_TEXT, seed=7) + inner[:200000] + bytes(100000) + inner
    for level, strategy in ((0, 0), (6, 2), (6, 0), (1, 3), (9, 1)):
        co = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
        s = co.compress(payload) + co.flush()
        res, out = _stream(ctx, s, refz.WRAP_GZIP, len(payload) + 16)
        assert res.status == 0 and out == payload and res.in_used == len(s), (level, strategy, msg(res.status), res.out_len)
    # the reference's statuses: output too small, truncated, damaged data, damaged trailer
    d = refz.gen(3000000, refz.GEN_MARKOV, seed=77)
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    s = co.compress(d) + co.flush()
    res, out = _stream(ctx, s, refz.WRAP_GZIP, 1000)
    assert msg(res.status) == "output buffer full" and res.out_len == len(d)
    for cut in (len(s) // 2, len(s) - 3, len(s) - 9):
        res, out = _stream(ctx, s[:cut], refz.WRAP_GZIP, len(d) + 16)
        one, res1 = ctx.inflate_host(s[:cut], [(0, cut, 0, len(d) + 16)], refz.WRAP_GZIP, 1)
        assert msg(res.status) == "truncated input" and res.status == res1[0].status and d.startswith(out), (cut, msg(res.status))
    for at, bit in ((len(s) // 3, 0x40), (len(s) // 2 + 7, 1), (40, 0x10)):
        bad = bytearray(s); bad[at] ^= bit
        res, out = _stream(ctx, bytes(bad), refz.WRAP_GZIP, len(d) + 16)
        one, res1 = ctx.inflate_host(bytes(bad), [(0, len(bad), 0, len(d) + 16)], refz.WRAP_GZIP, 1)
        assert res.status == res1[0].status and res.status != 0, (at, bit, msg(res.status), msg(res1[0].status))
    bad = bytearray(s); bad[-6] ^= 1
    res, out = _stream(ctx, bytes(bad), refz.WRAP_GZIP, len(d) + 16)
    assert msg(res.status) == "incorrect data check"
    bad = bytearray(s); bad[-2] ^= 1
    res, out = _stream(ctx, bytes(bad), refz.WRAP_GZIP, len(d) + 16)
    assert msg(res.status) == "incorrect length check"


def test_block_parallel_decode_sees_distances_too_far_back(ctx):
    """The counting pass of the chunk decode does not know what lies before a chunk, so it cannot see a distance that
    reaches before the start of the stream; the second pass does, and its verdict is checked before its match list is
    used (tools/fuzz_blocks.py found both orders wrong: "output buffer full" reported ahead of the reference's "invalid
    distance too far back", and a half-written match list resolved).  A raw stream compressed behind a preset
    dictionary, decoded without it, is such a stream from its first matches on and valid everywhere else."""
    import zlib
    dic = refz.gen(32768, refz.GEN_MARKOV, seed=5)
    d = dic[1000:30000] + refz.gen(3000000, refz.GEN_MARKOV, seed=6)
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, 0, dic)
    s = co.compress(d) + co.flush()
    for cap in (len(d) + 64, 100000, 0):
        res, out = _stream(ctx, s, refz.WRAP_RAW, cap)
        one, res1 = ctx.inflate_host(s, [(0, len(s), 0, cap)], refz.WRAP_RAW, 1, out_size=max(cap, 1))
        assert msg(res1[0].status) == "invalid distance too far back"
        assert res.status == res1[0].status, (cap, msg(res.status))
    # with the dictionary in place the same stream is fine (and takes the chunk path where it is offered one: members given
    # a dictionary are decoded by the batch kernels)
    out, res2 = ctx.inflate_host(s, [(0, len(s), 32768, len(d), 0, 0, 32768)], refz.WRAP_RAW, 1, out_size=32768 + len(d), prefill=dic)
    assert res2[0].status == 0 and out[32768:32768 + res2[0].out_len] == d


def test_gunzip_few_large_members(ctx):
    """The usual .gz file — one member, or a few, each of many MB without a flush point (gzip / the reference's gzwrite):
    zb200_gunzip_host takes them one after the other through the single-stream decoders (block-header chunks in parallel)
    instead of one team per member; same bytes, same member table, trailing bytes ignored, damage reported."""
    import gzip
    parts = [refz.gen(n, kind, seed=n) for n, kind in ((7000000, refz.GEN_MARKOV), (5000000, refz.GEN_MIXED), (3000000, refz.GEN_TEXT))]
    blobs = [gzip.compress(p, 6) for p in parts]
    whole, plain = b"".join(blobs), b"".join(parts)
    ctx.profile(True)
    r, st, out, olen, nm, tab = _gunzip(ctx, whole + b"\0\0trailing", len(plain) + 64, 8)
    prof = ctx.profile_read()
    ctx.profile(False)
    assert r == 0 and st == 0 and out == plain and nm == 3, (r, st, olen, nm)
    assert any("count" in k for k in prof), sorted(prof)      # the chunk kernels decoded them
    off_in = off_out = 0
    for m, b, p in zip(tab, blobs, parts):
        assert (m.in_off, m.in_len, m.out_off, m.out_cap) == (off_in, len(b), off_out, len(p))
        off_in += len(b); off_out += len(p)
    r, st, out, olen, nm, tab = _gunzip(ctx, blobs[0], len(parts[0]) + 64, 8)
    assert r == 0 and st == 0 and out == parts[0] and nm == 1
    bad = bytearray(whole); bad[len(blobs[0]) + len(blobs[1]) // 2] ^= 0x20     # damage inside the second member: the first is delivered
    r, st, out, olen, nm, tab = _gunzip(ctx, bytes(bad), len(plain) + 64, 8)
    assert r == 0 and st != 0 and nm == 1 and out == parts[0], (r, st, nm, len(out))
    r, st, out, olen, nm, tab = _gunzip(ctx, whole, 1000, 8)                    # too small an output buffer: the size needed
    assert r == zb.ERR_OUTPUT and olen == len(plain), (r, olen)
