"""The oracle (oracle/zoracle.c) pinned to the reference: against the committed
golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py) and, where oracle/_ref/libzref.so is present,
against the compiled reference itself on fresh seeded inputs."""
import base64
import hashlib
import random

import pytest

import refz


@pytest.fixture(scope="module")
def o():
    return refz.oracle()


def _data(golden_entry):
    return refz.gen(golden_entry["n"], golden_entry["kind"])


def test_generator_fingerprint(golden, o):
    for e in golden["streams"]:
        if "data_sha256" in e:
            d = _data(e)
            assert hashlib.sha256(d).hexdigest() == e["data_sha256"], e["input"]
            assert o.crc32(d) == e["crc32"] and o.adler32(d) == e["adler32"]


def test_checksum_known_answers(golden, o):
    cs = golden["checksums"]
    big = refz.gen(cs["data_n"], cs["data_kind"])
    for c in cs["cases"]:
        d = big[c["off"]:c["off"] + c["n"]]
        assert o.crc32(d) == c["crc32"], c
        assert o.adler32(d) == c["adler32"], c
        assert o.crc32(d, 0xdeadbeef) == c["crc32_seeded"], c
        assert o.adler32(d, 0x12345678 % (65521 << 16) | 5) == c["adler32_seeded"], c
    assert o.c_crc32(0, None, 0) == cs["null"]["crc32"] == 0
    assert o.c_adler32(0, None, 0) == cs["null"]["adler32"] == 1


def test_combine_known_answers(golden, o):
    for c in golden["checksums"]["combine"]:
        assert o.c_crc32_combine(0x12345678, 0x9abcdef0, c["len2"]) == c["crc"], c
        assert o.c_crc32_combine_gen(c["len2"]) == c["gen"], c
        assert o.c_crc32_combine_op(0x12345678, 0x9abcdef0, c["gen"]) == c["crc"], c
        assert o.c_adler32_combine(0x00c8012d, 0x11e60398, c["len2"]) == c["adler"], c


def test_combine_identity(o):
    rng = random.Random(7)
    d = refz.gen(50000, refz.GEN_BYTES)
    for _ in range(40):
        k = rng.randrange(0, len(d) + 1)
        a, b = d[:k], d[k:]
        assert o.c_crc32_combine(o.crc32(a), o.crc32(b), len(b)) == o.crc32(d)
        assert o.c_adler32_combine(o.adler32(a), o.adler32(b), len(b)) == o.adler32(d)


def test_deflate_golden_streams(golden, o):
    n = 0
    for e in golden["streams"]:
        if "sha256" not in e:
            continue
        s = o.deflate_stream(_data(e), e["level"], e["strategy"], e["wrap"], e["chunk"])
        assert len(s) == e["len"] and hashlib.sha256(s).hexdigest() == e["sha256"], e
        if "hex" in e:
            assert s.hex() == e["hex"]
        n += 1
    assert n > 250


def test_inflate_puff_vectors(golden, o):
    for v in golden["puff_vectors"]:
        raw = bytes.fromhex(v["hex"])
        err, msg, out, used = o.inflate_all(raw, refz.WRAP_RAW, cap=4096)
        if v["ret"] == refz.Z_STREAM_END:
            assert err == 0 and out.hex() == v["out_hex"] and used == v["total_in"], v
        elif v["ret"] == refz.Z_BUF_ERROR:
            assert msg == "truncated input", (v, msg)
        else:
            assert v["ret"] == refz.Z_DATA_ERROR and msg == v["msg"], (v, msg)


def test_inflate_zeros_raw(golden, o):
    z = golden["zeros_raw"]
    raw = base64.b64decode(z["b64"])
    err, msg, out, used = o.inflate_all(raw, refz.WRAP_RAW, cap=z["out_len"] + 10)
    assert err == 0 and len(out) == z["out_len"] == 1234567 and used == z["total_in"]
    assert o.crc32(out) == z["crc32"] == 0x70986ef5 and o.adler32(out) == z["adler32"] == 0xd7950001


def test_inflate_golden_streams(golden, o):
    for e in golden["streams"]:
        if "hex" not in e:
            continue
        err, msg, out, used = o.inflate_all(bytes.fromhex(e["hex"]), e["wrap"], cap=e["n"] + 16)
        assert err == 0 and out == _data(e) and used == e["len"], e


needs_ref = pytest.mark.skipif(not refz.have_ref(), reason="oracle/_ref/libzref.so not built")


@needs_ref
@pytest.mark.parametrize("kind", [refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_RANDOM, refz.GEN_MIXED, refz.GEN_BYTES])
def test_deflate_vs_reference(kind, o):
    r = refz.ref()
    d = refz.gen(400000, kind, seed=1234 + kind)
    for level in range(1, 10):
        for strategy in ((0, 1, 2, 3, 4) if level in (1, 6, 9) else (0,)):
            a = r.deflate_stream(d, level, strategy, refz.WRAP_ZLIB, 262144)
            b = o.deflate_stream(d, level, strategy, refz.WRAP_ZLIB, 262144)
            assert a == b, (kind, level, strategy, len(a), len(b))


@needs_ref
def test_deflate_vs_reference_edge_sizes(o):
    r = refz.ref()
    base = refz.gen(70000, refz.GEN_MARKOV, seed=99)
    for n in (0, 1, 2, 3, 4, 5, 257, 258, 259, 262, 263, 32767, 32768, 32769, 65535, 65536, 65537):
        for level in (1, 6):
            for wrap in (0, 1, 2):
                assert r.deflate_stream(base[:n], level, 0, wrap, 0) == o.deflate_stream(base[:n], level, 0, wrap, 0), (n, level, wrap)
    runs = bytes(300000)          # long zero run: 258-length matches, slides
    assert r.deflate_stream(runs, 6, 0, 1, 0) == o.deflate_stream(runs, 6, 0, 1, 0)
    assert r.deflate_stream(runs, 1, 3, 1, 0) == o.deflate_stream(runs, 1, 3, 1, 0)


@needs_ref
def test_inflate_vs_reference_and_puff(o):
    r = refz.ref()
    p = refz.puff()
    import ctypes as C
    for kind in (refz.GEN_TEXT, refz.GEN_MIXED, refz.GEN_RANDOM):
        d = refz.gen(300000, kind, seed=77)
        for level, strategy, wrap in ((1, 0, 0), (6, 0, 1), (9, 0, 2), (6, 4, 0), (6, 2, 1), (1, 3, 2)):
            s = r.deflate_stream(d, level, strategy, wrap, 100000)
            err, msg, out, used = o.inflate_all(s, wrap, cap=len(d) + 8)
            ret, rmsg, rout, rin = r.inflate_all(s, wrap, cap=len(d) + 8)
            assert err == 0 and out == d == rout and used == len(s) == rin and ret == refz.Z_STREAM_END
            if wrap == 0:   # puff only speaks raw deflate
                dl, sl = C.c_ulong(len(d) + 8), C.c_ulong(len(s))
                dst = C.create_string_buffer(len(d) + 8)
                assert p.puff(dst, C.byref(dl), s, C.byref(sl)) == 0 and dst.raw[:dl.value] == d


@needs_ref
def test_inflate_corruption_classes_match_reference(o):
    """Flip bits / truncate reference streams: the oracle must report the same
    error class (msg string) as the reference inflate, or the same output."""
    r = refz.ref()
    rng = random.Random(11)
    d = refz.gen(20000, refz.GEN_MIXED, seed=5)
    for wrap in (0, 1, 2):
        s = r.deflate_stream(d, 6, 0, wrap, 0)
        for trial in range(120):
            b = bytearray(s)
            if trial % 3 == 0:
                b = b[:rng.randrange(0, len(b))]
            else:
                i = rng.randrange(0, len(b))
                b[i] ^= 1 << rng.randrange(8)
            ret, rmsg, rout, rin = r.inflate_all(bytes(b), wrap, cap=len(d) + 64)
            err, msg, out, used = o.inflate_all(bytes(b), wrap, cap=len(d) + 64)
            if ret == refz.Z_STREAM_END:
                assert err == 0 and out == rout, (wrap, trial)
            elif ret == refz.Z_DATA_ERROR:
                assert msg == rmsg, (wrap, trial, msg, rmsg)
            else:
                assert msg in ("truncated input", "output buffer full"), (wrap, trial, ret, msg)
