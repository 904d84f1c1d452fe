"""Host replays of the product's __host__ __device__ cores (tests/emul/*.cpp compile
the same csrc/*.cuh headers the kernels use, with g++).  These validate the GF(2)
weighting, the inflate state machine and the deflate phases on the CPU before GPU
time is spent.  The replays are test binaries; nothing here ships."""
import ctypes as C
import os
import random
import subprocess
import zlib

import pytest

import refz

EMUL = os.path.join(refz.ROOT, "tests", "emul")


def _build(name):
    src, out = os.path.join(EMUL, name + ".cpp"), os.path.join(EMUL, "lib" + name + ".so")
    hdrs = [os.path.join(refz.ROOT, "zlib_wasm_b200", "csrc", f) for f in os.listdir(os.path.join(refz.ROOT, "zlib_wasm_b200", "csrc")) if f.endswith(("h", "cuh"))]
    if not os.path.exists(out) or any(os.path.getmtime(f) > os.path.getmtime(out) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-o", out, src])
    return C.CDLL(out)


def test_checksum_thread_decomposition():
    L = _build("ck_emul")
    L.emul_checksum.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    rng = random.Random(1)
    big = refz.gen(1500000, refz.GEN_BYTES)
    buf = C.create_string_buffer(big, len(big) + 64)
    base = C.addressof(buf)
    for trial in range(120):
        n = rng.choice([0, 1, 15, 16, 17, 31, 33, 100, 4095, 4096, 5000, 65536, 70001, rng.randrange(0, len(big))])
        off = rng.randrange(0, 17) if trial % 2 else 0
        n = min(n, len(big) - off)
        T, parts = rng.choice([32, 256, 1024]), rng.choice([1, 2, 3, 7, 148])
        ic, ia = rng.choice([0, 0xdeadbeef]), rng.choice([1, 0x12340005])
        c, a = C.c_uint32(), C.c_uint32()
        L.emul_checksum(base + off, n, T, parts, 3, ic, ia, C.byref(c), C.byref(a))
        d = big[off:off + n]
        assert (c.value, a.value) == (zlib.crc32(d, ic), zlib.adler32(d, ia)), (n, off, T, parts)


def _inflate_emul():
    L = _build("inf_emul")
    u64 = C.c_uint64
    L.emul_inflate.argtypes = [C.c_void_p, u64, C.c_void_p, u64, C.c_int, u64, u64] + [C.POINTER(u64)] * 2 + \
        [C.POINTER(C.c_uint32)] * 2 + [C.POINTER(u64)] * 2 + [C.POINTER(C.c_int)]

    def run(data, wrap, cap, off=0, resume=(0, 0), dst=None):
        buf = C.create_string_buffer(len(data) + 8 + off)
        C.memmove(C.addressof(buf) + off, data, len(data))
        dst = dst if dst is not None else C.create_string_buffer(max(cap, 1))
        iu, ol, cb, co = u64(), u64(), u64(), u64()
        ck, isz, kind = C.c_uint32(), C.c_uint32(), C.c_int()
        st = L.emul_inflate(C.addressof(buf) + off, len(data), dst, cap, wrap, resume[0], resume[1], C.byref(iu), C.byref(ol),
                            C.byref(ck), C.byref(isz), C.byref(cb), C.byref(co), C.byref(kind))
        return st, dst.raw[:ol.value], iu.value, ck.value, isz.value, cb.value, co.value, kind.value, dst
    return run


def test_inflate_state_machine(golden):
    run = _inflate_emul()
    o = refz.oracle()
    msgs = [o.c_inflate_msg(i).decode() for i in range(22)]
    for v in golden["puff_vectors"]:
        st, out, iu = run(bytes.fromhex(v["hex"]), 0, 4096)[:3]
        if v["ret"] == 1:
            assert st == 0 and out.hex() == v["out_hex"] and iu == v["total_in"]
        elif v["ret"] == -5:
            assert msgs[st] == "truncated input"
        else:
            assert msgs[st] == v["msg"]
    rng = random.Random(3)
    for kind in (refz.GEN_TEXT, refz.GEN_MIXED, refz.GEN_RANDOM):
        d = refz.gen(300000, kind, seed=42 + kind)
        for lvl, strat, wrap, chunk in ((1, 0, 0, 0), (6, 0, 1, 100000), (9, 0, 2, 0), (6, 4, 0, 65536), (6, 2, 1, 0), (1, 3, 2, 0)):
            s = o.deflate_stream(d, lvl, strat, wrap, chunk)
            st, out, iu, ck, isz = run(s, wrap, len(d) + 3, off=rng.randrange(0, 8))[:5]
            assert st == 0 and out == d and iu == len(s)
            if wrap == 2:
                assert ck == zlib.crc32(d) and isz == len(d)
    # resume at a block boundary after truncated input
    d = refz.gen(500000, refz.GEN_MARKOV, seed=9)
    s = o.deflate_stream(d, 6, 0, refz.WRAP_GZIP, 0)
    st, out, iu, ck, isz, cb, co, kind, dst = run(s[:len(s) // 2], 3, len(d) + 8)
    assert msgs[st] == "truncated input" and cb > 0 and out == d[:len(out)] and kind == 2
    st, out, iu = run(s, kind, len(d) + 8, resume=(cb, co), dst=dst)[:3]
    assert st == 0 and out == d and iu == len(s)


def _rounds_emul():
    L = _build("inf_emul")
    u64 = C.c_uint64
    L.emul_inflate_rounds.argtypes = [C.c_void_p, u64, C.c_void_p, u64, C.c_int, C.c_int] + [C.POINTER(u64)] * 2 + \
        [C.POINTER(C.c_uint32)] * 2 + [C.POINTER(u64)] * 2 + [C.POINTER(C.c_int)]
    L.emul_round_stats.argtypes = [C.POINTER(u64 * 7)]

    def run(data, wrap, cap, off=0, force_lg=-1):
        buf = C.create_string_buffer(len(data) + 8 + off)
        C.memmove(C.addressof(buf) + off, data, len(data))
        dst = C.create_string_buffer(max(cap, 1))
        iu, ol, cb, co = u64(), u64(), u64(), u64()
        ck, isz, kind = C.c_uint32(), C.c_uint32(), C.c_int()
        st = L.emul_inflate_rounds(C.addressof(buf) + off, len(data), dst, cap, wrap, force_lg, C.byref(iu), C.byref(ol),
                                   C.byref(ck), C.byref(isz), C.byref(cb), C.byref(co), C.byref(kind))
        return st, dst.raw[:ol.value], iu.value, ck.value, isz.value

    def stats():
        a = (u64 * 7)()
        L.emul_round_stats(C.byref(a))
        return dict(zip(("rounds", "fix_passes", "fix_lane_runs", "serial_returns", "matches", "dep_matches", "copy_passes"), a))
    return run, stats


def test_inflate_rounds(golden):
    """The self-synchronising rounds (zb_inflate_round.cuh) give the bytes, the consumed
    length and — through the serial fallback — the status of the serial state machine."""
    run, stats = _rounds_emul()
    serial = _inflate_emul()
    o = refz.oracle()
    msgs = [o.c_inflate_msg(i).decode() for i in range(22)]
    rng = random.Random(5)
    for kind in (refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_MIXED, refz.GEN_RANDOM):
        d = refz.gen(400000, kind, seed=77 + kind)
        for lvl, strat, wrap, chunk in ((1, 0, 0, 0), (6, 0, 1, 100000), (9, 0, 2, 0), (6, 4, 0, 65536), (6, 2, 1, 0), (1, 3, 2, 0), (6, 1, 2, 0)):
            s = o.deflate_stream(d, lvl, strat, wrap, chunk)
            for lg in (-1, 2, 3, 1000, 1003):                      # 1000 + k: a team of 128 lanes, S forced to 32 << (k - 1)
                st, out, iu, ck, isz = run(s, wrap, len(d) + rng.randrange(0, 4), off=rng.randrange(0, 8), force_lg=lg)
                assert st == 0 and out == d and iu == len(s), (kind, lvl, strat, wrap, lg, msgs[st])
                if wrap == 2:
                    assert ck == zlib.crc32(d) and isz == len(d)
    st = stats()
    assert st["rounds"] > 100 and st["fix_passes"] < 6 * st["rounds"], st      # incl. the forced S = 128 / 256 runs
    d = refz.gen(1 << 20, refz.GEN_MARKOV, seed=3)
    s = o.deflate_stream(d, 6, 0, 2, 0)
    assert run(s, 2, len(d))[1] == d
    st2 = stats()
    rounds, fixes = st2["rounds"] - st["rounds"], st2["fix_passes"] - st["fix_passes"]
    assert rounds > 50 and fixes < 2 * rounds, (rounds, fixes)                 # S = 1024: ~1.4 fix-up passes per round
    # small and degenerate members: below the round threshold everything runs on the serial path
    for n in (0, 1, 10, 600, 5000, 20000):
        d = refz.gen(n, refz.GEN_TEXT, seed=n)
        s = o.deflate_stream(d, 6, 0, 2, 0)
        st_, out, iu = run(s, 2, n)[:3]
        assert st_ == 0 and out == d and iu == len(s)
    # errors: same status and the same bytes before the error as the serial machine
    d = refz.gen(300000, refz.GEN_MARKOV, seed=5)
    s = bytearray(o.deflate_stream(d, 6, 0, 1, 0))
    for trial in range(40):
        bad = bytearray(s)
        where = rng.randrange(2, len(bad) - 4)
        bad[where] ^= 1 << rng.randrange(8)
        cap = len(d) + 64
        a = run(bytes(bad), 1, cap, force_lg=1000 if trial % 2 else -1)
        b = serial(bytes(bad), 1, cap)
        assert a[0] == b[0] and a[2] == b[2], (trial, where, msgs[a[0]], msgs[b[0]])
        if a[0] == 0:
            assert a[1] == b[1]
        else:
            n = min(len(a[1]), len(b[1]))
            assert len(a[1]) == len(b[1]) and a[1][:n] == b[1][:n]
    # truncated input and short output
    for cut in (len(s) // 3, len(s) - 5, len(s) - 1):
        a, b = run(bytes(s[:cut]), 1, len(d)), serial(bytes(s[:cut]), 1, len(d))
        assert a[0] == b[0] and msgs[a[0]] == "truncated input" and a[1] == b[1]
    a, b = run(bytes(s), 1, len(d) - 1000), serial(bytes(s), 1, len(d) - 1000)
    assert a[0] == b[0] and msgs[a[0]] == "output buffer full"
    for v in golden["puff_vectors"]:
        a, b = run(bytes.fromhex(v["hex"]), 0, 4096), serial(bytes.fromhex(v["hex"]), 0, 4096)
        assert a[0] == b[0] and a[1] == b[1]


def test_deflate_phases_byte_exact():
    L = _build("def_emul")
    L.emul_deflate_chunk.restype = C.c_long
    L.emul_deflate_chunk.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
    o = refz.oracle()
    for kind in range(5):
        for n in (0, 1, 3, 300, 70000, 262144):
            d = refz.gen(n, kind, seed=7 + kind)
            for level, strat in ((0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (6, 0), (6, 1), (6, 2), (6, 3), (6, 4), (9, 0)):
                for final in (0, 1):
                    cap = n + n // 8 + 1024
                    out, st = C.create_string_buffer(cap), (C.c_uint32 * 2)()
                    r = L.emul_deflate_chunk(d, n, level, strat, final, out, cap, st)
                    assert r >= 0, (r, kind, n, level, strat)
                    e = out.raw[:r]
                    rb = C.create_string_buffer(cap)
                    rn = o.c_deflate_chunk(d, n, level, strat, final, rb, cap)
                    if level == 0:                         # stored blocks of MAX_STORED bytes: the reference's bytes when it has room (deflate.c:1635-1815)
                        if refz.have_ref() and final:
                            assert e == refz.ref().deflate_stream(d, 0, 0, refz.WRAP_RAW, 0), (kind, n)
                        s = e if final else e + b"\x03\x00"
                        err, msg, back, used = o.inflate_all(s, 0, cap=n + 16)
                        assert err == 0 and back == d and len(e) <= n + 5 * (n // 65535 + 2) + 5
                    elif level >= 4 or strat in (2, 3):
                        assert e == rb.raw[:rn], (kind, n, level, strat, final)
                    else:
                        s = e if final else e + b"\x03\x00"
                        err, msg, back, used = o.inflate_all(s, 0, cap=n + 16)
                        assert err == 0 and back == d and len(e) <= 1.03 * rn + 8


def test_deflate_preset_dictionary_byte_exact():
    """A preset dictionary (deflate.c:550-632) = history ahead of the first chunk: the phases, entered at the
    dictionary's end, emit the bytes the reference emits after deflateSetDictionary (levels 4-9, Z_RLE,
    Z_HUFFMAN_ONLY); levels 1-3 decode with the reference's inflate + inflateSetDictionary."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    L = _build("def_emul")
    L.emul_deflate_chunk_dict.restype = C.c_long
    L.emul_deflate_chunk_dict.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_uint32)]
    ref = refz.ref()
    rng = random.Random(3)
    for kind in (refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_MIXED):
        base = refz.gen(400000, kind, seed=21 + kind)
        for dl, n in ((1, 5000), (2, 1), (3, 0), (100, 300), (262, 70000), (5000, 262144), (32768, 100000), (32767, 40000), (32768, 262144),
                      (40000, 90000)):
            # the dictionary shares material with the data so that matches reach into it
            dictionary = base[100000:100000 + dl]
            data = (base[100000 + dl // 2:100000 + dl // 2 + n // 2] + base[:n - n // 2])[:n]
            tail = dictionary[-32768:]
            for level, strat in ((1, 0), (3, 0), (4, 0), (6, 0), (6, 1), (6, 2), (6, 3), (9, 0), (0, 0)):
                want = ref.deflate_stream(data, level, strat, refz.WRAP_RAW, 0, dictionary=dictionary)
                joined = tail + data
                cap = len(joined) + len(joined) // 8 + 1024
                out, st = C.create_string_buffer(cap), (C.c_uint32 * 2)()
                r = L.emul_deflate_chunk_dict(joined, len(joined), len(tail), level, strat, 1, out, cap, st)
                assert r >= 0, (r, kind, dl, n, level, strat)
                got = out.raw[:r]
                if level >= 4 or strat in (2, 3) or level == 0:
                    assert got == want, (kind, dl, n, level, strat, len(got), len(want))
                else:
                    err, msg, back, used = ref.inflate_all(got, refz.WRAP_RAW, cap=n + 16, dictionary=dictionary)
                    assert err == 1 and back == data and len(got) <= 1.03 * len(want) + 8, (kind, dl, n, level, strat, err, msg)


def test_carried_history_chunks_host_replay():
    """zb200.h ZB200_CHUNK_CARRY on the host: chunk c of a call is the dictionary picture with the min(32 KiB, bytes before
    it) in front of it as history and a sync point (finish = 0) or the final block at its end (zb_deflate.cu chunk_lo /
    chunk_data / chunk_len say exactly that to the kernels).  The phases then emit, chunk by chunk, the bytes of the
    reference's deflateSetDictionary(previous 32 KiB) + deflate(Z_SYNC_FLUSH) — and the whole is one valid stream."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    L = _build("def_emul")
    L.emul_deflate_chunk_dict.restype = C.c_long
    L.emul_deflate_chunk_dict.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_uint32)]
    ref = refz.ref()
    for kind, n, chunk in ((refz.GEN_TEXT, 300000, 65536), (refz.GEN_MIXED, 200000, 20000), (refz.GEN_MARKOV, 150000, 100000)):
        d = refz.gen(n, kind, seed=50 + kind)
        for level, strat in ((6, 0), (9, 1), (4, 0), (6, 3), (1, 0)):
            got, want = [], []
            nch = (n + chunk - 1) // chunk
            for c in range(nch):
                pos = c * chunk
                hist, piece = d[max(0, pos - 32768):pos], d[pos:pos + chunk]
                last = c == nch - 1
                joined = hist + piece
                cap = len(joined) + len(joined) // 8 + 1024
                out, st = C.create_string_buffer(cap), (C.c_uint32 * 2)()
                r = L.emul_deflate_chunk_dict(joined, len(joined), len(hist), level, strat, 1 if last else 0, out, cap, st)
                assert r >= 0, (r, kind, c, level, strat)
                got.append(out.raw[:r])
                want.append(ref.deflate_stream(piece, level, strat, refz.WRAP_RAW, 0, dictionary=hist if hist else None,
                                               last_flush=refz.Z_FINISH if last else refz.Z_SYNC_FLUSH))
            if level >= 4:
                assert got == want, (kind, chunk, level, strat, [len(x) for x in got], [len(x) for x in want])
            err, msg, back, used = ref.inflate_all(b"".join(got), refz.WRAP_RAW, cap=n + 16)
            assert err == 1 and back == d and used == sum(len(x) for x in got), (kind, chunk, level, strat, msg)


def test_carried_history_chunks_golden():
    """The same claim pinned WITHOUT the compiled reference: tests/golden/carry_golden.json (made from the unmodified
    reference by tests/golden/make_carry_golden.py) holds length and SHA-256 of the reference's carried-chunk streams;
    the host replay of the device cores reproduces them."""
    import hashlib
    import json
    g = json.load(open(os.path.join(refz.ROOT, "tests", "golden", "carry_golden.json")))
    L = _build("def_emul")
    L.emul_deflate_chunk_dict.restype = C.c_long
    L.emul_deflate_chunk_dict.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_uint32)]
    assert len(g["cases"]) >= 12
    for e in g["cases"]:
        d = refz.gen(e["n"], e["kind"], seed=e["seed"])
        chunk, n = e["chunk"], e["n"]
        nch = max(1, (n + chunk - 1) // chunk)
        parts = []
        for c in range(nch):
            pos = c * chunk
            hist, piece = d[max(0, pos - 32768):pos], d[pos:pos + chunk]
            joined = hist + piece
            cap = len(joined) + len(joined) // 8 + 1024
            out, st = C.create_string_buffer(cap), (C.c_uint32 * 2)()
            r = L.emul_deflate_chunk_dict(joined, len(joined), len(hist), e["level"], e["strategy"], 1 if c == nch - 1 else 0, out, cap, st)
            assert r >= 0, (r, e)
            parts.append(out.raw[:r])
        s = b"".join(parts)
        assert [len(p) for p in parts] == e["chunk_lens"], (e["kind"], e["n"], e["chunk"], e["level"])
        assert len(s) == e["len"] and hashlib.sha256(s).hexdigest() == e["sha256"], (e["kind"], e["n"], e["chunk"], e["level"])
        if "hex" in e:
            assert s.hex() == e["hex"]
        assert zlib.decompress(s, -15) == d


def _bits_stream_of_31_bit_matches(n_matches):
    """stored block of 20000 bytes, then a fixed-Huffman block of matches that take 31 bits each (length code 284 +
    5 extra bits, distance code 28 + 13 extra bits): the pattern on which the lean symbol loop refills in the
    middle of every match, 13 matches in a row, without passing its safe-zone check."""
    rng = random.Random(3)
    raw = bytes(rng.randrange(256) for _ in range(20000))
    out = bytearray([0x00]) + (20000).to_bytes(2, "little") + (20000 ^ 0xffff).to_bytes(2, "little") + raw
    acc, nb = 0, 0

    def put(v, k):                                    # LSB-first field
        nonlocal acc, nb
        acc |= v << nb; nb += k

    def put_code(code, k):                            # Huffman code: most significant bit first
        for i in range(k - 1, -1, -1):
            put((code >> i) & 1, 1)
    put(0, 1); put(1, 2)                              # BFINAL = 0, fixed block
    expect = bytearray(raw)
    for j in range(n_matches):
        put_code(0xC0 + 4, 8); put(j % 31, 5)         # length symbol 284: 227 + extra
        put_code(28, 5); put((j * 37) % 3000, 13)     # distance symbol 28: 16385 + extra
        ln, dist = 227 + j % 31, 16385 + (j * 37) % 3000
        for _ in range(ln):
            expect.append(expect[-dist])
    while nb >= 8 or (nb and True):
        out.append(acc & 0xff); acc >>= 8; nb = max(0, nb - 8)
        if nb == 0:
            break
    return bytes(out), bytes(expect)


def test_truncated_input_never_reads_past_its_end():
    """A stream cut at any byte decodes to a prefix of the full output, whatever bytes happen to follow the cut in
    memory (the lean symbol loop once chained its unconditional mid-match refills past the end of the input)."""
    L = _build("inf_emul")
    u64 = C.c_uint64
    L.emul_inflate.argtypes = [C.c_void_p, u64, C.c_void_p, u64, C.c_int, u64, u64] + [C.POINTER(u64)] * 2 + \
        [C.POINTER(C.c_uint32)] * 2 + [C.POINTER(u64)] * 2 + [C.POINTER(C.c_int)]
    o = refz.oracle()
    rng = random.Random(7)
    synth, synth_out = _bits_stream_of_31_bit_matches(60)
    cases = [(synth, synth_out, list(range(20006, len(synth))))]
    for kind, level in ((refz.GEN_MARKOV, 6), (refz.GEN_MARKOV, 1)):
        d = refz.gen(150000, kind, seed=40 + kind)
        s = o.deflate_stream(d, level, 0, refz.WRAP_RAW, 0)
        cases.append((s, d, sorted(set([rng.randrange(1, len(s)) for _ in range(200)] + list(range(len(s) - 600, len(s)))))))
    for s, d, cuts in cases:
        cap = len(d) + 64
        last = 0
        kind = level = 0
        for k in cuts:
            outs = []
            for pad in (b"\x00", b"\xff", b"\x5a"):
                buf = C.create_string_buffer(s[:k] + pad * 32, k + 32)
                dst = C.create_string_buffer(cap)
                iu, ol, cb, co = u64(), u64(), u64(), u64()
                ck, isz, kd = C.c_uint32(), C.c_uint32(), C.c_int()
                st = L.emul_inflate(C.addressof(buf), k, dst, cap, 0, 0, 0, C.byref(iu), C.byref(ol), C.byref(ck), C.byref(isz),
                                    C.byref(cb), C.byref(co), C.byref(kd))
                outs.append((st, dst.raw[:ol.value]))
            assert outs[0] == outs[1] == outs[2], (kind, level, k)
            st, out = outs[0]
            assert st == 19 and d.startswith(out) and len(out) >= last, (kind, level, k, st, len(out), last)   # 19: truncated input
            last = len(out)


def test_truncated_input_rounds_replay():
    """The same property through the rounds (32 and 128 lanes) with the serial tail behind them: a cut stream gives a
    prefix of the output that does not depend on what follows the cut in memory."""
    L = _build("inf_emul")
    u64 = C.c_uint64
    L.emul_inflate_rounds.argtypes = [C.c_void_p, u64, C.c_void_p, u64, C.c_int, C.c_int] + [C.POINTER(u64)] * 2 + \
        [C.POINTER(C.c_uint32)] * 2 + [C.POINTER(u64)] * 2 + [C.POINTER(C.c_int)]
    o = refz.oracle()
    rng = random.Random(12)
    synth, synth_out = _bits_stream_of_31_bit_matches(300)
    d = refz.gen(300000, refz.GEN_MARKOV, seed=66)
    for s, want in ((synth, synth_out), (o.deflate_stream(d, 6, 0, refz.WRAP_RAW, 120000), d)):
        cap = len(want) + 64
        for k in sorted(set(rng.randrange(1, len(s)) for _ in range(120)) | set(range(len(s) - 80, len(s)))):
            for lanes in (-1, 1000):
                outs = []
                for pad in (b"\x00", b"\xff"):
                    buf = C.create_string_buffer(s[:k] + pad * 64, k + 64)
                    dst = C.create_string_buffer(cap)
                    iu, ol, cb, co = u64(), u64(), u64(), u64()
                    ck, isz, kd = C.c_uint32(), C.c_uint32(), C.c_int()
                    st = L.emul_inflate_rounds(C.addressof(buf), k, dst, cap, 0, lanes, C.byref(iu), C.byref(ol), C.byref(ck), C.byref(isz),
                                               C.byref(cb), C.byref(co), C.byref(kd))
                    outs.append((st, dst.raw[:ol.value]))
                assert outs[0] == outs[1] and outs[0][0] == 19 and want.startswith(outs[0][1]), (len(s), k, lanes)


def test_deflate_window_bits_and_mem_level_byte_exact():
    """deflateInit2_'s windowBits (w_size, MAX_DIST, slide period: deflate.c:440-443, deflate.h:298) and memLevel
    (hash_bits, lit_bufsize => symbols per block: deflate.c:444-455,512) — the phases, run with those values, emit the
    reference's bytes (levels 4-9, Z_RLE, Z_HUFFMAN_ONLY); levels 1-3 decode and stay within 3 %."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    L = _build("def_emul")
    L.emul_deflate_chunk_opts.restype = C.c_long
    L.emul_deflate_chunk_opts.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_size_t, C.POINTER(C.c_uint32)]
    ref = refz.ref()

    def ref_stream(d, level, strat, wbits, mem):
        strm = refz.ZStream()
        assert ref.deflateInit2_(C.byref(strm), level, 8, -wbits, mem, strat, ref.version, C.sizeof(refz.ZStream)) == 0
        cap = len(d) + len(d) // 4 + 4096
        src, dst = C.create_string_buffer(d, max(len(d), 1)), C.create_string_buffer(cap)
        strm.next_in, strm.avail_in = C.addressof(src), len(d)
        strm.next_out, strm.avail_out = C.addressof(dst), cap
        assert ref.deflate(C.byref(strm), refz.Z_FINISH) == refz.Z_STREAM_END
        out = dst.raw[:cap - strm.avail_out]
        ref.deflateEnd(C.byref(strm))
        return out

    for kind, n in ((refz.GEN_TEXT, 150000), (refz.GEN_MIXED, 262144), (refz.GEN_MARKOV, 70001), (refz.GEN_RANDOM, 40000)):
        d = refz.gen(n, kind, seed=31 + kind)
        for wbits, mem in ((15, 8), (9, 8), (12, 8), (15, 1), (15, 9), (10, 3), (14, 9), (9, 1), (13, 5)):
            for level, strat in ((1, 0), (4, 0), (6, 0), (6, 1), (9, 0), (6, 3), (6, 2)):
                want = ref_stream(d, level, strat, wbits, mem)
                cap = n + n // 4 + 4096
                out, st = C.create_string_buffer(cap), (C.c_uint32 * 2)()
                r = L.emul_deflate_chunk_opts(d, n, 0, level, strat, wbits, mem, 1, out, cap, st)
                assert r >= 0, (r, kind, wbits, mem, level, strat)
                got = out.raw[:r]
                if level >= 4 or strat in (2, 3):
                    assert got == want, (kind, wbits, mem, level, strat, len(got), len(want))
                else:
                    err, msg, back, used = ref.inflate_all(got, refz.WRAP_RAW, cap=n + 16)
                    assert err == 1 and back == d and len(got) <= 1.03 * len(want) + 8, (kind, wbits, mem, level, strat, err, msg)


def test_one_run_shared_by_many_ctas_replay():
    """zb_deflate.cu hands a long chunk's ordered phases to many CTAs: chain links drawn range by range behind w_size
    re-inserted positions, the lazy parse handed from CTA to CTA (cold settle, provisional link, settle against the
    predecessor's provisional end, compare with its true end).  The host replay does the same with loops: the matches must
    be the ones of one head table over the whole chunk, the symbols and blocks the serial parse's, the bytes the
    reference's one-shot stream."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    L = _build("def_emul")
    L.emul_deflate_chunk_opts.restype = C.c_long
    L.emul_deflate_chunk_opts.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_size_t, C.POINTER(C.c_uint32)]
    L.emul_set_multi.argtypes = [C.c_uint32, C.c_uint32]
    ref = refz.ref()
    try:
        for kind, n in ((refz.GEN_TEXT, 300000), (refz.GEN_MIXED, 524288 + 77), (refz.GEN_MARKOV, 131072)):
            d = refz.gen(n, kind, seed=61 + kind)
            for G, rng in ((2, 65536), (3, 40960), (7, 16384), (16, 131072)):
                for level, strat, wbits, mem in ((6, 0, 15, 8), (9, 0, 15, 8), (4, 1, 15, 8), (6, 4, 12, 8), (5, 0, 15, 3)):
                    if rng < (1 << wbits) // 4:
                        continue
                    L.emul_set_multi(G, rng)
                    cap = n + n // 4 + 4096
                    out, st = C.create_string_buffer(cap), (C.c_uint32 * 4)()
                    r = L.emul_deflate_chunk_opts(d, n, 0, level, strat, wbits, mem, 1, out, cap, st)
                    assert r >= 0, (r, kind, n, G, rng, level, strat, wbits, mem)
                    strm = refz.ZStream()
                    assert ref.deflateInit2_(C.byref(strm), level, 8, -wbits, mem, strat, ref.version, C.sizeof(refz.ZStream)) == 0
                    src, dst = C.create_string_buffer(d, n), C.create_string_buffer(cap)
                    strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), n, C.addressof(dst), cap
                    assert ref.deflate(C.byref(strm), refz.Z_FINISH) == refz.Z_STREAM_END
                    want = dst.raw[:cap - strm.avail_out]
                    ref.deflateEnd(C.byref(strm))
                    assert out.raw[:r] == want, (kind, n, G, rng, level, strat, wbits, mem, r, len(want))
    finally:
        L.emul_set_multi(0, 0)


def test_keyed_heap_tree_builder_equals_the_transliterated_one():
    """tree_build_fast (what dfl_tree_kernel runs: the active heap carries (Freq << 8 | depth) beside the node number)
    must reproduce trees.c:627-706 exactly — the tie-breaks of trees.c:499-501 decide the code lengths.  Random, flat,
    geometric (length-limit overflow, trees.c:573-612), all-equal and depth-wrapping histograms."""
    L = _build("def_emul")
    L.emul_tree_compare.restype = C.c_int
    L.emul_tree_compare.argtypes = [C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.c_uint32, C.c_int]
    rng = random.Random(11)

    def run(lf, df, byte_len=1 << 20, strategy=0):
        a = (C.c_uint16 * 286)(*[min(65535, x) for x in lf])
        b = (C.c_uint16 * 30)(*[min(65535, x) for x in df])
        return L.emul_tree_compare(a, b, byte_len, strategy)

    cases = 0
    for trial in range(400):
        kind = trial % 8
        if kind == 0:
            lf = [rng.randrange(0, 200) for _ in range(286)]; df = [rng.randrange(0, 100) for _ in range(30)]
        elif kind == 1:                                   # many equal frequencies: ties everywhere
            v = rng.randrange(1, 5)
            lf = [v if rng.random() < 0.8 else 0 for _ in range(286)]; df = [v] * 30
        elif kind == 2:                                   # geometric: overflows the 15-bit limit
            lf = [0] * 286
            for i in range(rng.randrange(17, 40)):
                lf[rng.randrange(286)] = min(65535, int(1.6 ** i) + 1)
            df = [min(65535, 2 ** i) for i in range(30)]
        elif kind == 3:                                   # Fibonacci: the deepest trees
            f = [1, 1]
            while len(f) < 24:
                f.append(f[-1] + f[-2])
            lf = [0] * 286
            for i, x in enumerate(f):
                lf[i * 11] = min(65535, x)
            df = [min(65535, x) for x in f] + [0] * 6
        elif kind == 4:                                   # one or two symbols only (trees.c:655-661 forces two codes)
            lf = [0] * 286; df = [0] * 30
            lf[rng.randrange(256)] = rng.randrange(1, 1000)
            if rng.random() < 0.5:
                df[rng.randrange(30)] = 3
        elif kind == 5:                                   # text-like: skewed literals, some lengths / distances
            lf = [int(1000 / (1 + abs(i - 101))) if 32 <= i < 127 else 0 for i in range(286)]
            for i in range(257, 286):
                lf[i] = rng.randrange(0, 300)
            df = [rng.randrange(0, 400) for _ in range(30)]
        elif kind == 6:                                   # all 286 symbols in use with equal counts
            lf = [57] * 286; df = [57] * 30
        else:                                             # large counts (ush sums wrap like the reference's)
            lf = [rng.randrange(0, 65535) if rng.random() < 0.1 else rng.randrange(0, 3) for _ in range(286)]
            df = [rng.randrange(0, 65535) for _ in range(30)]
        for strategy in (0, 4):
            for byte_len in (1 << 20, 10):
                assert run(lf, df, byte_len, strategy) == 0, (trial, kind, strategy, byte_len)
                cases += 1
    assert cases == 1600


def test_block_parallel_decode_replay():
    """zb_inflate_blocks.cuh on the host: the header test over every bit position of reference-made streams, the
    state machine's count / list chunk modes, chain, source pointers, pointer jumping, gather — the original bytes."""
    L = _build("inf_emul")
    u64 = C.c_uint64
    L.emul_inflate_blocks.argtypes = [C.c_void_p, u64, C.c_int, C.c_void_p, u64, C.POINTER(u64), C.POINTER(C.c_int), C.POINTER(u64),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(u64)]
    L.emul_inflate_blocks.restype = C.c_int
    cases = [(refz.GEN_TEXT, 700000, 6, 15), (refz.GEN_TEXT, 400000, 1, 31), (refz.GEN_MIXED, 900000, 9, -15),
             (refz.GEN_TEXT, 300000, 4, 15), (refz.GEN_MIXED, 500000, 6, 31)]
    for gen, n, level, wbits in cases:
        d = refz.gen(n, gen, seed=n + level)
        co = zlib.compressobj(level, zlib.DEFLATED, wbits)
        comp = co.compress(d) + co.flush()
        wrap = 0 if wbits < 0 else 2 if wbits > 15 else 1
        dst = C.create_string_buffer(n + 64)
        ol, st, iu, ck, isz = u64(), C.c_int(), u64(), C.c_uint32(), C.c_uint32()
        stats = (u64 * 4)()
        r = L.emul_inflate_blocks(comp, len(comp), wrap, dst, n + 64, C.byref(ol), C.byref(st), C.byref(iu), C.byref(ck), C.byref(isz), stats)
        assert r >= 2, (gen, n, level, wbits, r, list(stats))
        assert st.value == 0 and ol.value == n and dst.raw[:n] == d, (gen, n, level, wbits, list(stats))
        assert iu.value == len(comp)
        if wrap == 2:
            assert ck.value == zlib.crc32(d) and isz.value == n
        if wrap == 1:
            assert ck.value == zlib.adler32(d)
        assert stats[0] <= 8 * stats[1] + 2, list(stats)      # next to no false candidates (a stored block has up to 8 possible starts)
    # adversarial: real block headers that are NOT block starts — a compressed stream carried as DATA inside stored blocks
    # (level 0) and inside a literal-only stream (Z_HUFFMAN_ONLY) — are candidates the chain has to step over; streams of
    # stored blocks only, and a stream cut at every kind of place, fall back or deliver what the chain proves
    inner = zlib.compress(refz.gen(400000, refz.GEN_TEXT, seed=5), 6)
    payload = refz.gen(100000, refz.GEN_MARKOV, seed=6) + inner + refz.gen(200000, refz.GEN_TEXT, seed=7) + inner[:50000]
    for level, strategy in ((0, 0), (6, 2), (6, 0), (1, 3)):
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        comp = co.compress(payload) + co.flush()
        n = len(payload)
        dst = C.create_string_buffer(n + 64)
        ol, st, iu, ck, isz = u64(), C.c_int(), u64(), C.c_uint32(), C.c_uint32()
        stats = (u64 * 4)()
        r = L.emul_inflate_blocks(comp, len(comp), 1, dst, n + 64, C.byref(ol), C.byref(st), C.byref(iu), C.byref(ck), C.byref(isz), stats)
        assert r >= 0, (level, strategy, r, list(stats))
        if r >= 2:                                           # (fewer than two chunks: the caller takes the one-member path)
            assert st.value == 0 and ol.value == n and dst.raw[:n] == payload and iu.value == len(comp), (level, strategy, list(stats))
            assert ck.value == zlib.adler32(payload)


def test_deflate_fast_exact_chains_byte_exact():
    """zb_deflate.cuh fast_exact_chunk (ZB200_EXACT_FAST): deflate_fast with the reference's own parse-dependent hash chains
    (deflate.c:1824-1915: insertions only where the loop stands and inside matches of at most max_insert_length) — levels 1-3
    byte for byte the reference's streams, on every generator, ragged sizes, Z_FILTERED / Z_FIXED, window sizes and memLevels."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    L = _build("def_emul")
    L.emul_deflate_chunk_opts.restype = C.c_long
    L.emul_deflate_chunk_opts.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_size_t, C.POINTER(C.c_uint32)]
    L.emul_set_exact_fast(1)
    try:
        ref = refz.ref()
        ZS = C.sizeof(refz.ZStream)

        def reference(d, level, strat, wbits, mem, final):
            strm = refz.ZStream()
            assert ref.deflateInit2_(C.byref(strm), level, 8, -wbits, mem, strat, ref.version, ZS) == 0
            cap = len(d) + len(d) // 8 + 1024
            src, dst = C.create_string_buffer(d, max(len(d), 1)), C.create_string_buffer(cap)
            strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(d), C.addressof(dst), cap
            r = ref.deflate(C.byref(strm), refz.Z_FINISH if final else refz.Z_FULL_FLUSH)
            assert r == (1 if final else 0) and strm.avail_in == 0
            out = dst.raw[:cap - strm.avail_out]
            ref.deflateEnd(C.byref(strm))
            return out

        rng = random.Random(11)
        cases = 0
        for kind in (refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_RANDOM, refz.GEN_MIXED, refz.GEN_BYTES):
            for n in (0, 1, 2, 3, 4, 262, 263, 1000, 32768, 65274, 65275, 65536, 70000, 131072, 262144, 300001):
                d = refz.gen(n, kind, seed=70 + kind)
                for level in (1, 2, 3):
                    strat = rng.choice((0, 0, 1, 4))
                    wbits, mem = rng.choice(((15, 8), (15, 8), (15, 9), (12, 8), (9, 1), (14, 4), (15, 1)))
                    final = rng.random() < 0.7
                    want = reference(d, level, strat, wbits, mem, final)
                    cap = n + n // 8 + 1024
                    out, st = C.create_string_buffer(cap), (C.c_uint32 * 2)()
                    r = L.emul_deflate_chunk_opts(d, n, 0, level, strat, wbits, mem, 1 if final else 0, out, cap, st)
                    assert r >= 0, (r, kind, n, level, strat, wbits, mem, final)
                    assert out.raw[:r] == want, (kind, n, level, strat, wbits, mem, final, r, len(want))
                    cases += 1
        assert cases >= 200
    finally:
        L.emul_set_exact_fast(0)
