"""GPU parity of the boundary items added in round 2: deflateInit2_'s windowBits / memLevel honoured byte for byte,
deflatePrime / deflateUsed, inflatePrime / inflateSync / inflateSyncPoint, zalloc / zfree, the LFS *64 names, the five
remaining Boundary-B exports, and deflateBound / compressBound on incompressible input."""
import ctypes as C
import os
import random
import subprocess
import sys

import pytest

import refz
import zlib_wasm_b200 as zb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZS = C.sizeof(refz.ZStream)


@pytest.fixture(scope="module")
def z():
    return refz.ZlibBinding(zb.LIB_PATH, "")


def _deflate_all(lib, d, level, wbits, mem, strat, flush_every=0, prime=None):
    """One z_stream through `lib` (product or reference): Z_FULL_FLUSH every `flush_every` bytes, Z_FINISH at the end."""
    strm = refz.ZStream()
    assert lib.deflateInit2_(C.byref(strm), level, 8, wbits, mem, strat, lib.version, ZS) == 0
    if prime:
        fn = getattr(lib.lib, lib.prefix + "deflatePrime")
        fn.restype, fn.argtypes = C.c_int, [C.POINTER(refz.ZStream), C.c_int, C.c_int]
        assert fn(C.byref(strm), prime[0], prime[1]) == 0
    cap = len(d) + len(d) // 4 + 65536
    src, dst = C.create_string_buffer(d, max(len(d), 1)), C.create_string_buffer(cap)
    step = flush_every or max(len(d), 1)
    off = produced = 0
    while True:
        k = min(step, len(d) - off)
        last = off + k >= len(d)
        strm.next_in, strm.avail_in = C.addressof(src) + off, k
        strm.next_out, strm.avail_out = C.addressof(dst) + produced, cap - produced
        r = lib.deflate(C.byref(strm), refz.Z_FINISH if last else refz.Z_FULL_FLUSH)
        assert r in (0, 1), r
        produced = cap - strm.avail_out
        off += k
        if last:
            assert r == 1
            break
    used = C.c_int(-1)
    if hasattr(lib.lib, lib.prefix + "deflateUsed"):
        fn = getattr(lib.lib, lib.prefix + "deflateUsed")
        fn.restype, fn.argtypes = C.c_int, [C.POINTER(refz.ZStream), C.POINTER(C.c_int)]
        fn(C.byref(strm), C.byref(used))
    lib.deflateEnd(C.byref(strm))
    return dst.raw[:produced], used.value


def test_window_bits_and_mem_level_are_the_references(z):
    """deflate.c:440-455,1006: w_size / MAX_DIST / slide period, hash_bits, lit_bufsize and the CINFO of the zlib header.
    Levels 4-9 byte-identical with the reference's stream for the same settings and Z_FULL_FLUSH chunking; the
    reference's inflateInit2 with the SAME windowBits decodes every level."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    chunk = 262144
    for kind, n in ((refz.GEN_MIXED, 3 * chunk + 777), (refz.GEN_TEXT, 200000)):
        d = refz.gen(n, kind, seed=91 + kind)
        for wbits, mem in ((9, 8), (12, 8), (15, 1), (15, 9), (10, 3), (14, 9), (9, 1)):
            for level, strat in ((6, 0), (9, 0), (4, 1), (1, 0), (6, 3)):
                for wrapped in (wbits, -wbits, wbits + 16):
                    if wrapped != wbits and (level, strat) != (6, 0):
                        continue
                    got, _ = _deflate_all(z, d, level, wrapped, mem, strat, chunk)
                    want, _ = _deflate_all(ref, d, level, wrapped, mem, strat, chunk)
                    if level >= 4 or strat == 3:
                        assert got == want, (kind, wbits, mem, level, strat, wrapped, len(got), len(want))
                    else:
                        assert len(got) <= 1.03 * len(want) + 16
                    # the reference's decoder with the same (small) window takes it
                    strm = refz.ZStream()
                    assert ref.inflateInit2_(C.byref(strm), wrapped, ref.version, ZS) == 0
                    src, dst = C.create_string_buffer(got, len(got)), C.create_string_buffer(n + 16)
                    strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(got), C.addressof(dst), n + 16
                    assert ref.inflate(C.byref(strm), refz.Z_FINISH) == refz.Z_STREAM_END, (wbits, mem, level, strm.msg)
                    assert dst.raw[:n] == d
                    ref.inflateEnd(C.byref(strm))
            if wbits < 15:                                  # CINFO advertises the window that was asked for
                hdr, _ = _deflate_all(z, d[:1000], 6, wbits, mem, 0)
                assert hdr[0] >> 4 == wbits - 8 and (hdr[0] * 256 + hdr[1]) % 31 == 0


def test_deflate_prime_and_used(z):
    """deflate.c:731-757 / :723: bits ahead of the first block, the blocks follow at that bit offset (the reference's
    bytes), and deflateUsed reports the bits in use in the last byte."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    d = refz.gen(250000, refz.GEN_MARKOV, seed=5)              # (under one 256 KiB chunk: the library cuts longer calls itself)
    for bits, value in ((3, 5), (8, 0xa5), (11, 0x5a5), (16, 0xbeef), (1, 1)):
        for level in (6, 9, 4):
            got, used_g = _deflate_all(z, d, level, -15, 8, 0, 0, prime=(bits, value))
            want, used_r = _deflate_all(ref, d, level, -15, 8, 0, 0, prime=(bits, value))
            assert got == want, (bits, level, len(got), len(want))
            if used_r >= 0:
                assert used_g == used_r, (bits, level, used_g, used_r)
        got, _ = _deflate_all(z, d, 6, -15, 8, 0, 100000, prime=(bits, value))     # ... with flush points behind it
        want, _ = _deflate_all(ref, d, 6, -15, 8, 0, 100000, prime=(bits, value))
        assert got == want
    # the primed bits decode as part of the stream: prime an empty stored block header's first 3 bits? no — check by decode:
    # a stream primed with the 3 header bits of a non-final stored block of length 0 would need alignment; instead
    # verify with inflatePrime below.


def _bind(z, name, res, *args):
    fn = getattr(z.lib, z.prefix + name)
    fn.restype, fn.argtypes = res, list(args)
    return fn


def test_inflate_prime(z):
    """inflate.c:223-240: bits ahead of the first input byte — a raw stream entered at a bit offset (examples/zran.c)."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    d = refz.gen(200000, refz.GEN_TEXT, seed=12)
    s = ref.deflate_stream(d, 6, 0, refz.WRAP_RAW, 0)
    big = int.from_bytes(s, "little")
    for k in (3, 7, 8, 13, 16):
        value = big & ((1 << k) - 1)
        rest = (big >> k).to_bytes((len(s) * 8 - k + 7) // 8, "little")
        for lib in (z, ref):
            strm = refz.ZStream()
            assert lib.inflateInit2_(C.byref(strm), -15, lib.version, ZS) == 0
            prime = _bind(lib, "inflatePrime", C.c_int, C.POINTER(refz.ZStream), C.c_int, C.c_int)
            assert prime(C.byref(strm), k, value) == 0
            src, dst = C.create_string_buffer(rest, len(rest)), C.create_string_buffer(len(d) + 16)
            strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(rest), C.addressof(dst), len(d) + 16
            r = lib.inflate(C.byref(strm), refz.Z_FINISH)
            assert r == refz.Z_STREAM_END, (k, r, strm.msg)
            assert dst.raw[:len(d)] == d and strm.total_out == len(d)
            assert strm.total_in <= len(rest)
            lib.inflateEnd(C.byref(strm))
    strm = refz.ZStream()
    assert z.inflateInit2_(C.byref(strm), -15, z.version, ZS) == 0
    prime = _bind(z, "inflatePrime", C.c_int, C.POINTER(refz.ZStream), C.c_int, C.c_int)
    assert prime(C.byref(strm), 17, 0) == refz.Z_STREAM_ERROR and prime(C.byref(strm), -1, 0) == 0 and prime(C.byref(strm), 0, 0) == 0
    z.inflateEnd(C.byref(strm))


def test_inflate_sync_and_sync_point(z):
    """inflate.c:1375-1421: after a data error, skip to the next full flush point (00 00 FF FF) and go on decoding from
    there; the bytes after the point are the reference's.  inflate.c:1431: inflateSyncPoint."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    chunk = 100000
    d = refz.gen(4 * chunk, refz.GEN_MARKOV, seed=44)
    for wrap, wbits in ((refz.WRAP_ZLIB, 15), (refz.WRAP_RAW, -15), (refz.WRAP_GZIP, 31)):
        s = bytearray(ref.deflate_stream(d, 6, 0, wrap, chunk))
        marks = [i for i in range(len(s) - 3) if s[i:i + 4] == b"\x00\x00\xff\xff"]
        assert len(marks) >= 3
        s[marks[0] + 4] = 0x06                               # the second run opens with block type 3: "invalid block type"
        results = []
        for lib in (z, ref):
            strm = refz.ZStream()
            assert lib.inflateInit2_(C.byref(strm), wbits, lib.version, ZS) == 0
            sync = _bind(lib, "inflateSync", C.c_int, C.POINTER(refz.ZStream))
            src, dst = C.create_string_buffer(bytes(s), len(s)), C.create_string_buffer(len(d) + 64)
            cut = marks[1] + 4 + 10                          # first call: up to a little past the second marker
            strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), cut, C.addressof(dst), len(d) + 64
            r = lib.inflate(C.byref(strm), refz.Z_NO_FLUSH)
            while r == refz.Z_OK and strm.avail_in:
                r = lib.inflate(C.byref(strm), refz.Z_NO_FLUSH)
            assert r == refz.Z_DATA_ERROR, (wrap, r)
            first = strm.total_out
            # the caller goes on from wherever the library left next_in (the reference stops at the error, this library
            # has taken the whole slice): sync over what is left of the slice, then feed the rest of the file
            r = sync(C.byref(strm))
            assert r == refz.Z_OK, (wrap, lib.prefix, r)
            out2 = C.create_string_buffer(len(d) + 64)
            strm.next_out, strm.avail_out = C.addressof(out2), len(d) + 64
            t0 = strm.total_out
            r = lib.inflate(C.byref(strm), refz.Z_NO_FLUSH)
            assert r in (refz.Z_OK, refz.Z_BUF_ERROR), (wrap, lib.prefix, r, strm.msg)
            assert strm.avail_in == 0
            strm.next_in, strm.avail_in = C.addressof(src) + cut, len(s) - cut
            r = lib.inflate(C.byref(strm), refz.Z_FINISH)
            assert r == refz.Z_STREAM_END, (wrap, lib.prefix, r, strm.msg)
            results.append((first <= chunk * 2, out2.raw[:strm.total_out - t0]))
            lib.inflateEnd(C.byref(strm))
        assert results[0][1] == results[1][1] == d[2 * chunk:], wrap
    # no pattern at all: everything is consumed, Z_DATA_ERROR; no input: Z_BUF_ERROR
    strm = refz.ZStream()
    assert z.inflateInit2_(C.byref(strm), -15, z.version, ZS) == 0
    sync = _bind(z, "inflateSync", C.c_int, C.POINTER(refz.ZStream))
    assert sync(C.byref(strm)) == refz.Z_BUF_ERROR
    junk = C.create_string_buffer(b"\x01\x02\x03" * 100, 300)
    strm.next_in, strm.avail_in = C.addressof(junk), 300
    assert sync(C.byref(strm)) == refz.Z_DATA_ERROR and strm.avail_in == 0 and strm.total_in == 300
    z.inflateEnd(C.byref(strm))
    # inflateSyncPoint: input that stops between a sync flush's stored-block header and its LEN bytes (how PPP uses it)
    s = ref.deflate_stream(d[:50000], 6, 0, refz.WRAP_RAW, 0, chunk_flush=refz.Z_SYNC_FLUSH)
    s2 = ref.deflate_stream(d[:100000], 6, 0, refz.WRAP_RAW, 50000, chunk_flush=refz.Z_SYNC_FLUSH)
    m = s2.index(b"\x00\x00\xff\xff")
    for lib in (z, ref):
        sp = _bind(lib, "inflateSyncPoint", C.c_int, C.POINTER(refz.ZStream))
        strm = refz.ZStream()
        assert lib.inflateInit2_(C.byref(strm), -15, lib.version, ZS) == 0
        src, dst = C.create_string_buffer(s2, len(s2)), C.create_string_buffer(200000)
        strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), m, C.addressof(dst), 200000
        lib.inflate(C.byref(strm), refz.Z_NO_FLUSH)
        assert sp(C.byref(strm)) == 1, lib.prefix
        strm.next_in, strm.avail_in = C.addressof(src) + m, 7
        lib.inflate(C.byref(strm), refz.Z_NO_FLUSH)
        assert sp(C.byref(strm)) == 0, lib.prefix
        lib.inflateEnd(C.byref(strm))


def test_zalloc_zfree_are_honoured(z):
    """deflate.c:393-406 / inflate.c:187-199: the stream state comes from the caller's allocator, and every block it
    handed out comes back through zfree by the End call."""
    ALLOC = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_uint, C.c_uint)
    FREE = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
    libc = C.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes = C.c_void_p, [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    live, stats = set(), {"alloc": 0, "free": 0, "opaque": 0}

    def alloc(opaque, items, size):
        p = libc.malloc(items * size)
        live.add(p)
        stats["alloc"] += 1
        stats["opaque"] = opaque
        return p

    def free(opaque, p):
        assert p in live
        live.discard(p)
        stats["free"] += 1
        libc.free(p)

    a, f = ALLOC(alloc), FREE(free)
    d = refz.gen(100000, refz.GEN_TEXT, seed=8)
    strm = refz.ZStream()
    strm.zalloc, strm.zfree, strm.opaque = C.cast(a, C.c_void_p), C.cast(f, C.c_void_p), 0x1234
    assert z.deflateInit2_(C.byref(strm), 6, 8, 15, 8, 0, z.version, ZS) == 0
    assert stats["alloc"] >= 1 and stats["opaque"] == 0x1234
    src, dst = C.create_string_buffer(d, len(d)), C.create_string_buffer(len(d) + 1024)
    strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(d), C.addressof(dst), len(d) + 1024
    assert z.deflate(C.byref(strm), refz.Z_FINISH) == refz.Z_STREAM_END
    comp = dst.raw[:strm.total_out]
    assert z.deflateEnd(C.byref(strm)) == 0 and not live and stats["free"] == stats["alloc"]
    strm = refz.ZStream()
    strm.zalloc, strm.zfree, strm.opaque = C.cast(a, C.c_void_p), C.cast(f, C.c_void_p), 0x77
    assert z.inflateInit2_(C.byref(strm), 15, z.version, ZS) == 0 and live
    src, dst = C.create_string_buffer(comp, len(comp)), C.create_string_buffer(len(d))
    strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(comp), C.addressof(dst), len(d)
    assert z.inflate(C.byref(strm), refz.Z_FINISH) == refz.Z_STREAM_END and dst.raw == d
    assert z.inflateEnd(C.byref(strm)) == 0 and not live
    # zalloc == NULL: the library settles on its own pair and writes it back (deflate.c:393-406)
    strm = refz.ZStream()
    assert z.deflateInit2_(C.byref(strm), 6, 8, 15, 8, 0, z.version, ZS) == 0 and strm.zalloc and strm.zfree
    z.deflateEnd(C.byref(strm))


def test_lfs_names_and_remaining_exports(z):
    """zlib.map's *64 names and the Boundary-B exports of src/wasm_module_side.c, zlib_simd_compression.c,
    zlib_simd_optimized.c."""
    L = C.CDLL(zb.LIB_PATH, mode=C.RTLD_LOCAL)
    o = refz.oracle()
    d = refz.gen(300000, refz.GEN_TEXT, seed=3)
    for nm in ("crc32_combine64", "adler32_combine64"):
        getattr(L, nm).restype, getattr(L, nm).argtypes = C.c_ulong, [C.c_ulong, C.c_ulong, C.c_long]
    L.crc32_combine_gen64.restype, L.crc32_combine_gen64.argtypes = C.c_ulong, [C.c_long]
    assert L.crc32_combine64(o.crc32(d[:7]), o.crc32(d[7:]), len(d) - 7) == o.crc32(d)
    assert L.adler32_combine64(o.adler32(d[:7]), o.adler32(d[7:]), len(d) - 7) == o.adler32(d)
    assert L.crc32_combine_gen64(12345) == z.crc32_combine_gen(12345)
    for nm in ("gzopen64", "gzseek64", "gztell64", "gzoffset64", "gzvprintf", "inflateCodesUsed", "deflateUsed"):
        assert hasattr(L, nm), nm
    # Boundary B
    L.zlib_compress_simd_buffer.restype = C.c_int
    L.zlib_compress_simd_buffer.argtypes = [C.c_char_p, C.c_ulong, C.c_void_p, C.POINTER(C.c_ulong), C.c_int]
    cap = z.compressBound(len(d))
    out, ol = C.create_string_buffer(cap), C.c_ulong(cap)
    assert L.zlib_compress_simd_buffer(d, len(d), out, C.byref(ol), 6) == 0
    back, bl = C.create_string_buffer(len(d)), C.c_ulong(len(d))
    assert z.uncompress(back, C.byref(bl), out.raw[:ol.value], ol.value) == 0 and back.raw == d
    L.zlib_crc32_simd.restype, L.zlib_crc32_simd.argtypes = C.c_ulong, [C.c_ulong, C.c_char_p, C.c_uint]
    assert L.zlib_crc32_simd(0, d, len(d)) == o.crc32(d)
    L.zlib_adler32_simd.restype, L.zlib_adler32_simd.argtypes = C.c_uint, [C.c_uint, C.c_char_p, C.c_size_t]
    assert L.zlib_adler32_simd(1, d, len(d)) == o.adler32(d)
    L.zlib_benchmark_simd_compression.restype = C.c_double
    L.zlib_benchmark_simd_compression.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
    assert L.zlib_benchmark_simd_compression(d, len(d), 2) > 0 and L.zlib_benchmark_simd_compression(None, 0, 1) == -1.0
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    L.zlib_simd_analysis.restype = None
    L.zlib_simd_analysis.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.zlib_simd_analysis(d, len(d), C.byref(a), C.byref(b), C.byref(c))
    assert 2.0 < a.value < 5.0 and b.value > 0 and c.value == 1.0
    L.zlib_simd_performance_analysis.restype = None
    L.zlib_simd_performance_analysis.argtypes = L.zlib_simd_analysis.argtypes
    L.zlib_simd_performance_analysis(d, len(d), C.byref(a), C.byref(b), C.byref(c))
    assert a.value > 0 and b.value > 0 and c.value > 0


BOUND_SCRIPT = r"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(%(root)r, "tests")); sys.path.insert(0, %(root)r)
import refz, zlib_wasm_b200 as zb
z = refz.ZlibBinding(zb.LIB_PATH, "")
ZS = C.sizeof(refz.ZStream)
for n in (0, 1, 1000, 100000, 5 * 262144 + 17, 16 << 20):
    d = refz.gen(n, refz.GEN_RANDOM, seed=n + 1)
    # compress() into exactly compressBound() bytes
    cap = z.compressBound(n)
    dst, dl = C.create_string_buffer(cap), C.c_ulong(cap)
    assert z.compress2(dst, C.byref(dl), d, n, 6) == 0, ("compress2", n, cap)
    for level, strat, mem in ((6, 2, 8), (6, 3, 8), (1, 0, 8), (9, 0, 8), (6, 2, 1), (6, 0, 9)):
        strm = refz.ZStream()
        assert z.deflateInit2_(C.byref(strm), level, 8, 15, mem, strat, z.version, ZS) == 0
        bound = z.deflateBound(C.byref(strm), n)
        src, out = C.create_string_buffer(d, max(n, 1)), C.create_string_buffer(bound)
        strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), n, C.addressof(out), bound
        r = z.deflate(C.byref(strm), refz.Z_FINISH)
        assert r == refz.Z_STREAM_END, ("deflateBound", n, level, strat, mem, bound, r)
        z.deflateEnd(C.byref(strm))
print("ok")
"""


@pytest.mark.parametrize("chunk", ["", "100000", "1024"])
def test_bounds_hold_for_incompressible_input(chunk):
    """compress.c:72 / deflate.c:842: compress() into a compressBound() buffer and deflate(Z_FINISH) into a deflateBound()
    buffer always succeed — random bytes, Z_HUFFMAN_ONLY / Z_RLE (literal-only blocks), memLevel 1 (127-symbol blocks),
    and chunk sizes that are no multiple of the block size ($ZB200_CHUNK is read once per process: subprocess)."""
    env = dict(os.environ)
    if chunk:
        env["ZB200_CHUNK"] = chunk
    else:
        env.pop("ZB200_CHUNK", None)
    p = subprocess.run([sys.executable, "-c", BOUND_SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-2000:]


@pytest.mark.parametrize("wbits", [15, 31, -15])
def test_streaming_inflate_memory_is_bounded(z, wbits):
    """A long stream WITHOUT flush points (the reference's one-shot compress2 / deflate(Z_FINISH): one run of blocks) fed in
    slices: the one-member path re-bases at block boundaries (32 KiB of history, as inflate.c:368-412 keeps), so what the
    stream holds stays small, and output, trailer verdict and total counts are the reference's."""
    if not refz.have_refpool():
        pytest.skip("oracle/_ref/librefpool.so not built")
    rp = refz.refpool()
    n = 48 << 20
    d = refz.gen(n, refz.GEN_MARKOV, seed=99)
    cap = n // 2 + (1 << 20)
    comp = C.create_string_buffer(cap)
    clen, sec = C.c_size_t(0), C.c_double(0)
    assert rp.rp_compress2_whole(d, n, 6, 0, wbits, 8, comp, cap, C.byref(clen), C.byref(sec)) == 0
    s = comp.raw[:clen.value]
    assert s.count(b"\x00\x00\xff\xff") < 50                 # no flush points (only chance occurrences)
    for slice_in, slice_out in ((1 << 16, 1 << 20), (3 << 20, 1 << 17)):
        strm = refz.ZStream()
        assert z.inflateInit2_(C.byref(strm), wbits, z.version, ZS) == 0
        src, dst = C.create_string_buffer(s, len(s)), C.create_string_buffer(n + 64)
        fed = produced = 0
        r = 0
        while r == 0 or r == refz.Z_BUF_ERROR:
            if strm.avail_in == 0 and fed < len(s):
                k = min(slice_in, len(s) - fed)
                strm.next_in, strm.avail_in = C.addressof(src) + fed, k
                fed += k
            room = min(slice_out, n + 64 - produced)
            strm.next_out, strm.avail_out = C.addressof(dst) + produced, room
            r = z.inflate(C.byref(strm), refz.Z_NO_FLUSH)
            produced += room - strm.avail_out
            if r == refz.Z_BUF_ERROR and strm.avail_in == 0 and fed >= len(s) and room == strm.avail_out:
                break
        assert r == refz.Z_STREAM_END, (wbits, slice_in, r, strm.msg, produced)
        assert produced == n == strm.total_out and strm.total_in == len(s)
        assert dst.raw[:n] == d
        z.inflateEnd(C.byref(strm))
    # a damaged trailer is still caught after any number of re-bases
    if wbits != -15:
        bad = bytearray(s)
        bad[-5 if wbits == 31 else -1] ^= 1
        ret, msg, out, tin = z.inflate_all(bytes(bad), {15: refz.WRAP_ZLIB, 31: refz.WRAP_GZIP}[wbits], cap=n + 64, in_slice=1 << 18)
        assert ret == refz.Z_DATA_ERROR and msg == "incorrect data check", (ret, msg)


def test_one_shot_calls_emit_the_references_own_stream(z):
    """compress.c:22-59: the reference's compress2() emits ONE run of blocks, and so does this library — levels 4-9 byte for
    byte the reference's compress2 output, on all five generators (here at config C1's size; long inputs in
    test_one_shot_calls_of_any_size_are_the_references_stream); levels 1-3 decode and stay within 3 % (theirs is one run up
    to 1 MiB, Z_FULL_FLUSH chunks beyond)."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    for kind in (refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_RANDOM, refz.GEN_MIXED, refz.GEN_BYTES):
        for n in (1 << 20, 700001, 262145):
            d = refz.gen(n, kind, seed=7 * kind + 1)
            for level in (1, 4, 6, 9):
                cap = z.compressBound(n)
                a, al = C.create_string_buffer(cap), C.c_ulong(cap)
                b, bl = C.create_string_buffer(cap + 1024), C.c_ulong(cap + 1024)
                assert z.compress2(a, C.byref(al), d, n, level) == 0
                assert ref.compress2(b, C.byref(bl), d, n, level) == 0
                if level >= 4:
                    assert a.raw[:al.value] == b.raw[:bl.value], (kind, n, level, al.value, bl.value)
                else:
                    assert al.value <= 1.03 * bl.value + 16
                    back, kl = C.create_string_buffer(n), C.c_ulong(n)
                    assert ref.uncompress(back, C.byref(kl), a.raw[:al.value], al.value) == 0 and back.raw == d
    # the streaming API, fed in slices with no flush until Z_FINISH, is the same stream (zlib.h:253-262)
    d = refz.gen(900000, refz.GEN_MARKOV, seed=77)
    for wrap in (refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP):
        assert z.deflate_stream(d, 6, 0, wrap, chunk=0, in_slice=10000, out_slice=4096) == ref.deflate_stream(d, 6, 0, wrap, 0)


def _run(exe, args, stdin=None, timeout=300):
    return subprocess.run([exe] + args, input=stdin, capture_output=True, timeout=timeout)


def test_reference_block_level_examples_linked_against_product(tmp_path):
    """inflate(Z_BLOCK) + strm->data_type (zlib.h:540-560, inflate.c:824-826,1248-1251), inflatePrime, inflateSetDictionary,
    deflatePrime: the reference's own block-level tools — examples/zran.c (random access through an index of block
    boundaries), gzjoin.c (members spliced at the bit level) and gzappend.c — compiled against the reference's zlib.h and
    linked to libzb200.so, untouched, give the results of the same tools linked to the reference."""
    import gzip
    bins = os.path.join(ROOT, "tests", "_bin")
    refs = os.path.join(ROOT, "oracle", "_ref")
    need = [os.path.join(bins, n + "_b200") for n in ("zran", "gzjoin", "gzappend")] + [os.path.join(refs, n) for n in ("zran", "gzjoin", "gzappend")]
    if not all(os.path.exists(p) for p in need) or not refz.have_ref():
        pytest.skip("the example binaries were not prebuilt (needs /root/reference at build time)")
    ref = refz.ref()
    d = refz.gen(6 << 20, refz.GEN_MARKOV, seed=61)
    # ---- zran: index of access points ~1 MiB apart, then 16 KiB extracted at several offsets
    files = {"a.gz": ref.deflate_stream(d, 6, 0, refz.WRAP_GZIP, 0), "a.zz": ref.deflate_stream(d, 9, 0, refz.WRAP_ZLIB, 0),
             "a.raw": ref.deflate_stream(d[:3 << 20], 6, 0, refz.WRAP_RAW, 0),
             "two.gz": ref.deflate_stream(d[:2 << 20], 6, 0, refz.WRAP_GZIP, 0) + ref.deflate_stream(d[2 << 20:], 1, 0, refz.WRAP_GZIP, 0)}
    for name, blob in files.items():
        p = str(tmp_path / name)
        open(p, "wb").write(blob)
        n = (3 << 20) if name == "a.raw" else len(d)
        for off in (0, 1, 1234567, n - 20000, n - 5, (n * 2 + 2) // 3):
            a = _run(need[0], [p, str(off)])
            b = _run(need[3], [p, str(off)])
            assert a.returncode == b.returncode == 0, (name, off, a.stderr, b.stderr)
            assert a.stdout == b.stdout == d[off:off + 16384][:max(0, n - off)], (name, off, len(a.stdout), len(b.stdout))
            assert a.stderr == b.stderr, (name, off, a.stderr, b.stderr)       # same number of access points, same byte counts
    # ---- gzjoin: the members' deflate streams spliced into ONE member at the bit level (no recompression)
    parts = [d[:700000], d[700000:700001], d[700001:2500000], d[2500000:2600000]]
    names = []
    for i, part in enumerate(parts):
        p = str(tmp_path / ("p%d.gz" % i))
        open(p, "wb").write(ref.deflate_stream(part, (6, 9, 1, 4)[i], 0, refz.WRAP_GZIP, 0))
        names.append(p)
    a = _run(need[1], names)
    b = _run(need[4], names)
    assert a.returncode == b.returncode == 0, (a.stderr, b.stderr)
    assert a.stdout == b.stdout, (len(a.stdout), len(b.stdout))
    assert gzip.decompress(b.stdout) == b"".join(parts)
    # ---- gzappend: scan to the last block (Z_BLOCK), then deflate behind it (deflatePrime + deflateSetDictionary)
    more = d[3 << 20:(3 << 20) + 600000]
    for tool, tag in ((need[2], "x"), (need[5], "y")):
        p = str(tmp_path / ("app_%s.gz" % tag))
        open(p, "wb").write(ref.deflate_stream(d[:900000], 6, 0, refz.WRAP_GZIP, 0))
        q = str(tmp_path / ("more_%s.bin" % tag))
        open(q, "wb").write(more)
        r = _run(tool, ["-6", p, q])
        assert r.returncode == 0, r.stderr
    xa, ya = open(str(tmp_path / "app_x.gz"), "rb").read(), open(str(tmp_path / "app_y.gz"), "rb").read()
    assert gzip.decompress(xa) == d[:900000] + more
    assert xa == ya


def _deflate_whole(lib, d, level, wbits, mem, strat):
    """deflateInit2 + one deflate(Z_FINISH) over the whole input: (stream bytes, adler)."""
    strm = refz.ZStream()
    assert lib.deflateInit2_(C.byref(strm), level, 8, wbits, mem, strat, lib.version, ZS) == 0
    cap = lib.deflateBound(C.byref(strm), len(d)) + 64
    src, dst = C.create_string_buffer(d, max(len(d), 1)), C.create_string_buffer(cap)
    strm.next_in, strm.avail_in, strm.next_out, strm.avail_out = C.addressof(src), len(d), C.addressof(dst), cap
    assert lib.deflate(C.byref(strm), refz.Z_FINISH) == refz.Z_STREAM_END
    out = dst.raw[:cap - strm.avail_out]
    lib.deflateEnd(C.byref(strm))
    return out


def test_one_shot_calls_of_any_size_are_the_references_stream(z):
    """compress.c:22-59 / deflate(Z_FINISH): ONE run of blocks however long the input.  Levels 4-9 emit it by default up to
    the engine's largest chunk (1 GiB): the ordered phases of the long chunk are shared by many CTAs — chain ranges behind
    32 KiB of re-inserted history, the lazy parse handed from CTA to CTA (zb_deflate.cu dfl_parse_multi_kernel), block
    ends from a scan.  Byte for byte the reference's stream: five generators at 64 MiB, a 300 MiB input, levels 4 / 6 / 9,
    Z_FILTERED / Z_FIXED, small windows and memLevels, sizes around the CTA ranges."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    n = 64 << 20
    for kind in (refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_RANDOM, refz.GEN_MIXED, refz.GEN_BYTES):
        d = refz.gen(n, kind, seed=40 + kind)
        cap = z.compressBound(n)
        a, al = C.create_string_buffer(cap), C.c_ulong(cap)
        b, bl = C.create_string_buffer(cap), C.c_ulong(cap)
        assert z.compress2(a, C.byref(al), d, n, 6) == 0
        assert ref.compress2(b, C.byref(bl), d, n, 6) == 0
        assert al.value == bl.value and a.raw[:al.value] == b.raw[:bl.value], (kind, al.value, bl.value)
    d = refz.gen(9000000, refz.GEN_MIXED, seed=51)
    for n in (524288, 524289, 786432 + 5, 2 * 262144 * 3 + 131, 9000000):
        for level, wbits, mem, strat in ((4, 15, 8, 0), (9, 15, 8, 0), (6, 15, 8, 1), (6, 15, 8, 4), (6, -12, 8, 0), (6, 31, 9, 0),
                                         (5, 10, 3, 0), (8, 15, 1, 0), (7, -15, 8, 1)):
            if n == 9000000 and level == 9:
                continue
            got, want = _deflate_whole(z, d[:n], level, wbits, mem, strat), _deflate_whole(ref, d[:n], level, wbits, mem, strat)
            assert got == want, (n, level, wbits, mem, strat, len(got), len(want))
    n = 300 << 20
    d = refz.gen(n, refz.GEN_MARKOV, seed=52)
    cap = z.compressBound(n)
    a, al = C.create_string_buffer(cap), C.c_ulong(cap)
    b, bl = C.create_string_buffer(cap), C.c_ulong(cap)
    assert z.compress2(a, C.byref(al), d, n, 4) == 0
    assert ref.compress2(b, C.byref(bl), d, n, 4) == 0
    assert al.value == bl.value and a.raw[:al.value] == b.raw[:bl.value], (al.value, bl.value)


CARRY_SCRIPT = r"""
import ctypes as C, os, sys, zlib
sys.path.insert(0, os.path.join(%(root)r, "tests")); sys.path.insert(0, %(root)r)
import refz, zlib_wasm_b200 as zb
z = refz.ZlibBinding(zb.LIB_PATH, "")
ZS = C.sizeof(refz.ZStream)
d = refz.gen(3000000, refz.GEN_MARKOV, seed=41)
# compress2 at level 1: beyond 1 MiB the greedy levels are cut into $ZB200_CHUNK chunks — carried, they see each other
for level in (1, 3, 6):
    cap = z.compressBound(len(d))
    dst, dl = C.create_string_buffer(cap), C.c_ulong(cap)
    assert z.compress2(dst, C.byref(dl), d, len(d), level) == 0
    s = dst.raw[:dl.value]
    assert zlib.decompress(s) == d
    print("size", level, len(s))
# a stream with flushes, a small window, a dictionary: every combination still the caller's bytes back
for wbits, mem, dic in ((15, 8, None), (-15, 8, d[:20000]), (9, 8, None), (12, 1, None), (31, 9, None)):
    strm = refz.ZStream()
    assert z.deflateInit2_(C.byref(strm), 2, 8, wbits, mem, 0, z.version, ZS) == 0
    if dic is not None:
        assert z.deflateSetDictionary(C.byref(strm), dic, len(dic)) == 0
    cap = len(d) + len(d) // 4 + 65536
    src, dst = C.create_string_buffer(d, len(d)), C.create_string_buffer(cap)
    off = produced = 0
    step = 700001
    while off < len(d):
        k = min(step, len(d) - off)
        strm.next_in, strm.avail_in = C.addressof(src) + off, k
        strm.next_out, strm.avail_out = C.addressof(dst) + produced, cap - produced
        fl = refz.Z_FINISH if off + k >= len(d) else (refz.Z_SYNC_FLUSH, refz.Z_FULL_FLUSH, refz.Z_NO_FLUSH)[(off // step) %% 3]
        r = z.deflate(C.byref(strm), fl)
        assert r in (0, 1), r
        produced = cap - strm.avail_out
        off += k
    assert r == 1
    z.deflateEnd(C.byref(strm))
    s = dst.raw[:produced]
    o = zlib.decompressobj(wbits if wbits != 9 else 15, zdict=dic) if dic is not None else zlib.decompressobj(wbits if wbits != 9 else 15)
    assert o.decompress(s) == d and o.eof, (wbits, mem)
# no flushes at all, $ZB200_STREAM_HOLD_MIB=1: the library cuts the stream itself every MiB — carried, the pieces see each other
for level in (6, 1):
    strm = refz.ZStream()
    assert z.deflateInit2_(C.byref(strm), level, 8, 15, 8, 0, z.version, ZS) == 0
    cap = len(d) + len(d) // 4 + 65536
    src, dst = C.create_string_buffer(d, len(d)), C.create_string_buffer(cap)
    off = produced = 0
    while True:
        k = min(300000, len(d) - off)
        strm.next_in, strm.avail_in = C.addressof(src) + off, k
        strm.next_out, strm.avail_out = C.addressof(dst) + produced, cap - produced
        off += k
        r = z.deflate(C.byref(strm), refz.Z_FINISH if off >= len(d) else refz.Z_NO_FLUSH)
        produced = cap - strm.avail_out
        assert r in (0, 1), r
        if r == 1:
            break
    z.deflateEnd(C.byref(strm))
    assert zlib.decompress(dst.raw[:produced]) == d
    print("size", 100 + level, produced)
print("ok")
"""


def test_chunk_carry_knob_at_the_zlib_surface():
    """$ZB200_CHUNK_CARRY=1 (read once per process: subprocess): wherever zlib.h calls are cut into chunks, the chunks are
    compressed behind the 32 KiB before them (zb200.h ZB200_CHUNK_CARRY).  Streams stay valid for every window size,
    memLevel, dictionary and flush mixture (Python's zlib = an independent decoder), and the chunked levels get smaller."""
    sizes = {}
    for knob in ("0", "1"):
        env = dict(os.environ)
        env["ZB200_CHUNK_CARRY"] = knob
        env["ZB200_CHUNK"] = "65536"
        env["ZB200_STREAM_HOLD_MIB"] = "1"
        p = subprocess.run([sys.executable, "-c", CARRY_SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-2000:]
        sizes[knob] = {int(l.split()[1]): int(l.split()[2]) for l in p.stdout.splitlines() if l.startswith("size")}
    assert sizes["1"][1] < sizes["0"][1] and sizes["1"][3] < sizes["0"][3]
    assert sizes["1"][6] == sizes["0"][6]                     # (one run of blocks either way: nothing is cut at level 6)
    assert sizes["1"][106] < sizes["0"][106] and sizes["1"][101] < sizes["0"][101]   # the library's own cuts keep the window, too
