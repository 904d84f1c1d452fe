"""GPU parity of the drop-in boundary: the zlib.h API and the src/wasm_module.c
exports served by libzb200.so, driven exactly like the reference's callers
(one-shot compress2/uncompress, zpipe-style sliced deflate()/inflate(), the
WASM shim's buffer and streaming functions), compared with the reference."""
import ctypes as C
import os
import random
import subprocess

import pytest

import refz
import zlib_wasm_b200 as zb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def z():
    return refz.ZlibBinding(zb.LIB_PATH, "")


def ref_or_oracle_stream(d, level, strategy, wrap, chunk):
    return (refz.ref() if refz.have_ref() else refz.oracle()).deflate_stream(d, level, strategy, wrap, chunk)


def lib_chunking(n):
    """How the zlib layer cuts one call's input: up to 1 MiB ($ZB200_SINGLE_RUN_MAX) not at all — the reference's own
    one-shot stream — beyond that into 256 KiB Z_FULL_FLUSH runs ($ZB200_CHUNK)."""
    return 0 if n <= (1 << 20) else 262144


def test_version_and_init_errors(z):
    assert z.zlibVersion() == b"1.3.1.1-motley"
    s = refz.ZStream()
    assert z.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, b"2.0", C.sizeof(refz.ZStream)) == refz.Z_VERSION_ERROR
    assert z.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, z.version, C.sizeof(refz.ZStream) - 8) == refz.Z_VERSION_ERROR
    for bad in ((10, 8, 15, 8, 0), (6, 7, 15, 8, 0), (6, 8, 7, 8, 0), (6, 8, 15, 0, 0), (6, 8, 15, 10, 0), (6, 8, 15, 8, 5), (6, 8, -16, 8, 0)):
        assert z.deflateInit2_(C.byref(s), *bad, z.version, C.sizeof(refz.ZStream)) == refz.Z_STREAM_ERROR, bad
    assert z.inflateInit2_(C.byref(s), 48, z.version, C.sizeof(refz.ZStream)) == refz.Z_STREAM_ERROR
    assert z.zError(-3) == b"data error" and z.zError(1) == b"stream end"


def test_checksums_like_zlib_h(z):
    o = refz.oracle()
    d = refz.gen(1 << 20, refz.GEN_TEXT)
    assert z.crc32(0, None, 0) == 0 and z.adler32(0, None, 0) == 1
    assert z.crc32(0, d, len(d)) == o.crc32(d) and z.adler32(1, d, len(d)) == o.adler32(d)
    c = a = None
    c, a = 0, 1
    for k in range(0, len(d), 100001):      # running use, zlib.h:1711-1768
        piece = d[k:k + 100001]
        c, a = z.crc32(c, piece, len(piece)), z.adler32(a, piece, len(piece))
    assert (c, a) == (o.crc32(d), o.adler32(d))
    assert z.crc32_combine(o.crc32(d[:5]), o.crc32(d[5:]), len(d) - 5) == o.crc32(d)
    assert z.adler32_combine(o.adler32(d[:5]), o.adler32(d[5:]), len(d) - 5) == o.adler32(d)


def test_compress2_uncompress_config_c1(z):
    """BASELINE config C1: level 6 round trip + checksums on a 1 MiB text buffer."""
    d = refz.gen(1 << 20, refz.GEN_TEXT)
    bound = z.compressBound(len(d))
    dst = C.create_string_buffer(bound)
    dl = C.c_ulong(bound)
    assert z.compress2(dst, C.byref(dl), d, len(d), 6) == refz.Z_OK
    s = dst.raw[:dl.value]
    assert s == ref_or_oracle_stream(d, 6, 0, refz.WRAP_ZLIB, 0)           # byte-identical with the reference's own compress2()
    if refz.have_ref():                                                     # the reference's own uncompress takes it
        rdst, rdl = C.create_string_buffer(bound), C.c_ulong(bound)
        assert refz.ref().compress2(rdst, C.byref(rdl), d, len(d), 6) == refz.Z_OK and rdst.raw[:rdl.value] == s
        r = refz.ref()
        back = C.create_string_buffer(len(d))
        bl = C.c_ulong(len(d))
        assert r.uncompress(back, C.byref(bl), s, len(s)) == refz.Z_OK and back.raw == d
    back = C.create_string_buffer(len(d))
    bl = C.c_ulong(len(d))
    assert z.uncompress(back, C.byref(bl), s, len(s)) == refz.Z_OK and bl.value == len(d) and back.raw == d
    # error mapping of uncompr.c:76-79
    bl = C.c_ulong(len(d) - 1)
    assert z.uncompress(back, C.byref(bl), s, len(s)) == refz.Z_BUF_ERROR
    bl = C.c_ulong(len(d))
    assert z.uncompress(back, C.byref(bl), s, len(s) - 9) == refz.Z_DATA_ERROR
    bad = bytearray(s); bad[len(s) // 2] ^= 0x10
    bl = C.c_ulong(len(d))
    assert z.uncompress(back, C.byref(bl), bytes(bad), len(s)) == refz.Z_DATA_ERROR
    tiny = C.c_ulong(8)
    assert z.compress2(dst, C.byref(tiny), d, len(d), 6) == refz.Z_BUF_ERROR


@pytest.mark.parametrize("wrap", [refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP])
def test_streaming_zpipe_style(z, wrap):
    """16 KiB in / 16 KiB out slicing (examples/zpipe.c:57-82,96-150)."""
    d = refz.gen(1500000, refz.GEN_MIXED, seed=12)
    for level in (1, 6):
        s = z.deflate_stream(d, level, 0, wrap, chunk=0, in_slice=16384, out_slice=16384)
        if level >= 4:                                       # the reference's own stream for the same calls: one run of blocks
            assert s == ref_or_oracle_stream(d, level, 0, wrap, 0)
        ret, msg, out, tin = z.inflate_all(s, wrap, cap=len(d) + 64, in_slice=16384, out_slice=16384)
        assert ret == refz.Z_STREAM_END and out == d and tin == len(s), (ret, msg)
        if refz.have_ref():
            ret, msg, out, tin = refz.ref().inflate_all(s, wrap, cap=len(d) + 64)
            assert ret == refz.Z_STREAM_END and out == d
    # explicit full flushes at odd places stay decodable and resumable
    s = z.deflate_stream(d, 6, 0, wrap, chunk=333333, in_slice=50000, out_slice=7777)
    ret, msg, out, tin = z.inflate_all(s, wrap, cap=len(d) + 64, in_slice=1000, out_slice=100000)
    assert ret == refz.Z_STREAM_END and out == d


def test_inflate_errors_and_trailing_bytes(z):
    d = refz.gen(300000, refz.GEN_MARKOV, seed=3)
    s = ref_or_oracle_stream(d, 6, 0, refz.WRAP_GZIP, 0)
    ret, msg, out, tin = z.inflate_all(s + b"TRAILING", refz.WRAP_GZIP, cap=len(d) + 64, in_slice=4096)
    assert ret == refz.Z_STREAM_END and out == d and tin == len(s)          # next_in stops right after the member
    bad = bytearray(s); bad[-5] ^= 1
    ret, msg, out, tin = z.inflate_all(bytes(bad), refz.WRAP_GZIP, cap=len(d) + 64)
    assert ret == refz.Z_DATA_ERROR and msg == "incorrect data check"
    ret, msg, out, tin = z.inflate_all(s[:len(s) // 2], refz.WRAP_GZIP, cap=len(d) + 64)
    assert ret == refz.Z_BUF_ERROR and out == d[:len(out)]
    ret, msg, out, tin = z.inflate_all(b"\x78\x9c" + b"\x07" * 10, refz.WRAP_ZLIB, cap=64)
    assert ret == refz.Z_DATA_ERROR and msg == "invalid block type"


def test_wasm_module_exports():
    L = C.CDLL(zb.LIB_PATH, mode=C.RTLD_LOCAL)
    o = refz.oracle()
    d = refz.gen(200000, refz.GEN_TEXT, seed=8)
    L.zlib_compress_bound.restype = C.c_ulong
    L.zlib_compress_bound.argtypes = [C.c_ulong]
    L.zlib_crc32.restype = L.zlib_adler32.restype = C.c_ulong
    L.zlib_crc32.argtypes = L.zlib_adler32.argtypes = [C.c_ulong, C.c_char_p, C.c_uint]
    L.zlib_get_version.restype = C.c_char_p
    L.zlib_compress_buffer.argtypes = [C.c_char_p, C.c_ulong, C.c_void_p, C.POINTER(C.c_ulong), C.c_int]
    L.zlib_decompress_buffer.argtypes = [C.c_void_p, C.c_ulong, C.c_void_p, C.POINTER(C.c_ulong)]
    assert L.zlib_get_version() == b"1.3.1.1-motley" and L.zlib_has_simd() == 0 and L.zlib_simd_capabilities() == 1
    assert L.zlib_crc32(0, d, len(d)) == o.crc32(d) and L.zlib_adler32(1, d, len(d)) == o.adler32(d)
    cap = L.zlib_compress_bound(len(d))
    dst = C.create_string_buffer(cap)
    dl = C.c_ulong(cap)
    assert L.zlib_compress_buffer(d, len(d), dst, C.byref(dl), 99) == 0     # bad level -> default (wasm_module.c:41-43)
    assert dst.raw[:dl.value] == ref_or_oracle_stream(d, 6, 0, refz.WRAP_ZLIB, lib_chunking(len(d)))
    assert L.zlib_compress_buffer(None, len(d), dst, C.byref(dl), 6) == refz.Z_STREAM_ERROR
    assert L.zlib_compress_buffer(d, 0, dst, C.byref(dl), 6) == refz.Z_STREAM_ERROR
    back = C.create_string_buffer(len(d))
    bl = C.c_ulong(len(d))
    assert L.zlib_decompress_buffer(dst, dl.value, back, C.byref(bl)) == 0 and back.raw == d
    # streaming context (wasm_module.c:153-288)
    L.zlib_deflate_init.restype = L.zlib_inflate_init.restype = C.c_void_p
    L.zlib_deflate_process.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, C.c_void_p, C.c_uint, C.c_int]
    L.zlib_inflate_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
    L.zlib_deflate_end.argtypes = L.zlib_inflate_end.argtypes = [C.c_void_p]
    L.zlib_stream_total_out.restype = C.c_ulong
    L.zlib_stream_total_out.argtypes = [C.c_void_p]
    h = L.zlib_deflate_init(6, 15, 8, 0)
    assert h
    out = C.create_string_buffer(cap)
    assert L.zlib_deflate_process(h, d, len(d), out, cap, refz.Z_FINISH) == refz.Z_STREAM_END
    n = L.zlib_stream_total_out(h)
    L.zlib_deflate_end(h)
    assert out.raw[:n] == dst.raw[:dl.value]
    h = L.zlib_inflate_init(15)
    back = C.create_string_buffer(len(d))
    assert L.zlib_inflate_process(h, out, n, back, len(d)) == refz.Z_STREAM_END and back.raw == d
    L.zlib_inflate_end(h)
    # raw-deflate "simd" one-shot (src/zlib_simd_optimized.c:354-383)
    L.zlib_compress_simd.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_int]
    ol = C.c_size_t(cap)
    assert L.zlib_compress_simd(d, len(d), out, C.byref(ol), 6) == 0
    assert out.raw[:ol.value] == ref_or_oracle_stream(d, 6, 0, refz.WRAP_RAW, lib_chunking(len(d)))


def test_reference_zpipe_linked_against_product(tmp_path):
    """The reference's examples/zpipe.c, compiled against the reference's zlib.h
    and linked to libzb200.so (prebuilt where /root/reference exists)."""
    exe = os.path.join(refz.ROOT, "tests", "_bin", "zpipe_b200")
    if not os.path.exists(exe):
        pytest.skip("zpipe_b200 was not prebuilt (needs /root/reference at build time)")
    d = refz.gen(1 << 20, refz.GEN_TEXT)
    comp = subprocess.run([exe], input=d, capture_output=True, timeout=120)
    assert comp.returncode == 0, comp.stderr
    assert comp.stdout == ref_or_oracle_stream(d, 6, 0, refz.WRAP_ZLIB, 0)        # zpipe uses Z_DEFAULT_COMPRESSION; 1 MiB = one run
    back = subprocess.run([exe, "-d"], input=comp.stdout, capture_output=True, timeout=120)
    assert back.returncode == 0 and back.stdout == d
    if refz.have_ref():
        zp = os.path.join(refz.ROOT, "oracle", "_ref", "zpipe")
        assert subprocess.run([zp, "-d"], input=comp.stdout, capture_output=True, timeout=120).stdout == d


@pytest.mark.parametrize("wrap", [refz.WRAP_RAW, refz.WRAP_ZLIB])
def test_preset_dictionary(z, wrap):
    """deflateSetDictionary / inflateSetDictionary (deflate.c:550-632, inflate.c:1278-1312): the reference's bytes
    at levels >= 4 on the same chunking, and either side decodes the other's stream."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    base = refz.gen(1200000, refz.GEN_TEXT, seed=41)
    for dl, n in ((100, 5000), (32768, 600000), (50000, 262144 + 1000), (1000, 0)):
        dic = base[:dl]
        d = base[dl // 2:dl // 2 + n]
        for level in (1, 6, 9):
            s = z.deflate_stream(d, level, 0, wrap, chunk=0, dictionary=dic)
            want = ref.deflate_stream(d, level, 0, wrap, lib_chunking(len(d)), dictionary=dic)
            if level >= 4:
                assert s == want, (dl, n, level, len(s), len(want))
            else:
                assert len(s) <= 1.03 * len(want) + 8
            ret, m, out, tin = ref.inflate_all(s, wrap, cap=n + 64, dictionary=dic)
            assert ret == refz.Z_STREAM_END and out == d, (dl, n, level, ret, m)
            ret, m, out, tin = z.inflate_all(want, wrap, cap=n + 64, dictionary=dic, in_slice=50000)
            assert ret == refz.Z_STREAM_END and out == d and tin == len(want), (dl, n, level, ret, m)
    if wrap == refz.WRAP_ZLIB:
        s = ref.deflate_stream(base[:100000], 6, 0, wrap, 0, dictionary=base[200000:230000])
        ret, m, out, tin = z.inflate_all(s, wrap, cap=100064)                       # nobody supplies it
        assert ret == refz.Z_NEED_DICT and out == b""
        ret, m, out, tin = z.inflate_all(s, wrap, cap=100064, dictionary=base[1:30001])   # the wrong one: inflate.c:1293-1296
        assert ret == refz.Z_DATA_ERROR
    with pytest.raises(RuntimeError):                                               # gzip takes no dictionary: deflate.c:562
        z.deflate_stream(base[:1000], 6, 0, refz.WRAP_GZIP, 0, dictionary=base[:100])


def _gz(z):
    L = z.lib
    if not hasattr(z, "_gz_bound"):
        L.gzopen.restype = C.c_void_p; L.gzopen.argtypes = [C.c_char_p, C.c_char_p]
        L.gzread.restype = C.c_int; L.gzread.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
        L.gzwrite.restype = C.c_int; L.gzwrite.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
        L.gzclose.restype = C.c_int; L.gzclose.argtypes = [C.c_void_p]
        L.gzeof.restype = C.c_int; L.gzeof.argtypes = [C.c_void_p]
        L.gzdirect.restype = C.c_int; L.gzdirect.argtypes = [C.c_void_p]
        L.gzgets.restype = C.c_void_p; L.gzgets.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.gzputs.restype = C.c_int; L.gzputs.argtypes = [C.c_void_p, C.c_char_p]
        L.gzgetc.restype = C.c_int; L.gzgetc.argtypes = [C.c_void_p]
        L.gzungetc.restype = C.c_int; L.gzungetc.argtypes = [C.c_int, C.c_void_p]
        L.gzseek.restype = C.c_long; L.gzseek.argtypes = [C.c_void_p, C.c_long, C.c_int]
        L.gztell.restype = C.c_long; L.gztell.argtypes = [C.c_void_p]
        L.gzflush.restype = C.c_int; L.gzflush.argtypes = [C.c_void_p, C.c_int]
        L.gzerror.restype = C.c_char_p; L.gzerror.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.gzrewind.restype = C.c_int; L.gzrewind.argtypes = [C.c_void_p]
        z._gz_bound = True
    return L


def _gz_read_all(L, path, step=1 << 20):
    f = L.gzopen(path.encode(), b"rb")
    assert f
    buf, out = C.create_string_buffer(step), b""
    while True:
        k = L.gzread(f, buf, step)
        assert k >= 0, k
        out += buf.raw[:k]
        if k < step:
            break
    return f, out


def test_gz_file_layer(z, tmp_path):
    """gzopen / gzwrite / gzread / gzgets / gzseek / gzclose (gzlib.c, gzread.c, gzwrite.c): files written here are
    read by Python's gzip module (an independent reader) and equal the reference's gzip stream; multi-member files
    written by others — concatenation, append mode — are read back whole; plain files pass through."""
    import gzip
    L = _gz(z)
    d = refz.gen(3000000, refz.GEN_TEXT, seed=77)
    p1 = str(tmp_path / "a.gz")
    f = L.gzopen(p1.encode(), b"wb6")
    assert f
    for off in range(0, len(d), 700001):
        piece = d[off:off + 700001]
        assert L.gzwrite(f, piece, len(piece)) == len(piece)
    assert L.gztell(f) == len(d) and L.gzclose(f) == 0
    raw = open(p1, "rb").read()
    assert gzip.decompress(raw) == d
    assert raw == ref_or_oracle_stream(d, 6, 0, refz.WRAP_GZIP, 0)   # the bytes the reference's gzwrite / gzclose leave
    # append mode adds a member; another writer's members follow; everything is read back as one stream
    f = L.gzopen(p1.encode(), b"ab9")
    assert L.gzputs(f, b"appended line\n") == 14 and L.gzclose(f) == 0
    more = [refz.gen(n, refz.GEN_MARKOV, seed=n) for n in (0, 10, 99999, 400000)]
    with open(p1, "ab") as fh:
        for m in more:
            fh.write(gzip.compress(m, 6))
    want = d + b"appended line\n" + b"".join(more)
    f, out = _gz_read_all(L, p1)
    assert out == want and L.gzeof(f) == 1 and L.gzdirect(f) == 0
    # seeking, single characters, lines
    assert L.gzseek(f, len(d), 0) == len(d) and L.gzeof(f) == 0
    line = C.create_string_buffer(64)
    assert L.gzgets(f, line, 64) and line.value == b"appended line\n"
    assert L.gzseek(f, -5, 1) == len(d) + 9 and L.gzgetc(f) == ord("l") and L.gzungetc(ord("L"), f) == ord("L")
    assert L.gzgetc(f) == ord("L") and L.gztell(f) == len(d) + 10
    assert L.gzrewind(f) == 0 and L.gzgetc(f) == d[0]
    assert L.gzclose(f) == 0
    # a plain file passes through (gzread.c:gz_look direct mode); "T" writes one
    p2 = str(tmp_path / "plain.txt")
    f = L.gzopen(p2.encode(), b"wT")
    assert L.gzwrite(f, d[:5000], 5000) == 5000 and L.gzclose(f) == 0
    assert open(p2, "rb").read() == d[:5000]
    f, out = _gz_read_all(L, p2)
    assert out == d[:5000] and L.gzdirect(f) == 1 and L.gzclose(f) == 0
    # a file cut short: what is there is delivered, the error says so (gzread.c:gz_decomp)
    p3 = str(tmp_path / "cut.gz")
    open(p3, "wb").write(raw[:len(raw) // 2])
    f, out = _gz_read_all(L, p3)
    err = C.c_int(0)
    m = L.gzerror(f, C.byref(err))
    assert out == b"" or d.startswith(out)
    assert err.value == refz.Z_BUF_ERROR and m == b"unexpected end of file" and L.gzclose(f) == refz.Z_BUF_ERROR
    assert not L.gzopen(p1.encode(), b"r+") and not L.gzopen(str(tmp_path / "nope.gz").encode(), b"rb")


@pytest.mark.parametrize("wrap", [refz.WRAP_RAW, refz.WRAP_GZIP])
def test_sync_flush_keeps_history(z, wrap):
    """Z_SYNC_FLUSH / Z_PARTIAL_FLUSH do not reset the window (deflate.c:1211-1218): small messages flushed one by
    one compress against what went before.  While the stream fits the window the bytes are the reference's."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    d = refz.gen(30000, refz.GEN_TEXT, seed=51)
    for level in (6, 9):
        s = z.deflate_stream(d, level, 0, wrap, chunk=3000, chunk_flush=refz.Z_SYNC_FLUSH)
        assert s == ref.deflate_stream(d, level, 0, wrap, chunk=3000, chunk_flush=refz.Z_SYNC_FLUSH), level
    d = refz.gen(2000000, refz.GEN_MARKOV, seed=52)
    for level, fl in ((1, refz.Z_SYNC_FLUSH), (6, refz.Z_SYNC_FLUSH), (6, refz.Z_PARTIAL_FLUSH)):
        s = z.deflate_stream(d, level, 0, wrap, chunk=4096, chunk_flush=fl)
        want = ref.deflate_stream(d, level, 0, wrap, chunk=4096, chunk_flush=refz.Z_SYNC_FLUSH)
        full = ref.deflate_stream(d, level, 0, wrap, chunk=4096, chunk_flush=refz.Z_FULL_FLUSH)
        assert len(s) <= 1.03 * len(want) and len(s) < 0.9 * len(full), (level, len(s), len(want), len(full))
        ret, m, out, tin = ref.inflate_all(s, wrap, cap=len(d) + 64)
        assert ret == refz.Z_STREAM_END and out == d
        ret, m, out, tin = z.inflate_all(s, wrap, cap=len(d) + 64, in_slice=300000)
        assert ret == refz.Z_STREAM_END and out == d
    # a full flush does reset it: the chunks decode on their own (zlib.h:286-291)
    s = z.deflate_stream(d[:100000], 6, 0, refz.WRAP_RAW, chunk=20000, chunk_flush=refz.Z_FULL_FLUSH)
    assert s == ref.deflate_stream(d[:100000], 6, 0, refz.WRAP_RAW, chunk=20000, chunk_flush=refz.Z_FULL_FLUSH)


class GzHeader(C.Structure):
    _fields_ = [("text", C.c_int), ("time", C.c_ulong), ("xflags", C.c_int), ("os", C.c_int),
                ("extra", C.c_void_p), ("extra_len", C.c_uint), ("extra_max", C.c_uint),
                ("name", C.c_void_p), ("name_max", C.c_uint), ("comment", C.c_void_p), ("comm_max", C.c_uint),
                ("hcrc", C.c_int), ("done", C.c_int)]


def _advanced(zb_):
    for name, res, args in (("deflateSetHeader", C.c_int, [C.POINTER(refz.ZStream), C.POINTER(GzHeader)]),
                            ("inflateGetHeader", C.c_int, [C.POINTER(refz.ZStream), C.POINTER(GzHeader)]),
                            ("deflatePending", C.c_int, [C.POINTER(refz.ZStream), C.POINTER(C.c_uint), C.POINTER(C.c_int)]),
                            ("deflateCopy", C.c_int, [C.POINTER(refz.ZStream), C.POINTER(refz.ZStream)]),
                            ("inflateCopy", C.c_int, [C.POINTER(refz.ZStream), C.POINTER(refz.ZStream)]),
                            ("inflateGetDictionary", C.c_int, [C.POINTER(refz.ZStream), C.c_void_p, C.POINTER(C.c_uint)]),
                            ("deflateGetDictionary", C.c_int, [C.POINTER(refz.ZStream), C.c_void_p, C.POINTER(C.c_uint)]),
                            ("inflateValidate", C.c_int, [C.POINTER(refz.ZStream), C.c_int]),
                            ("inflateMark", C.c_long, [C.POINTER(refz.ZStream)]),
                            ("zlibCompileFlags", C.c_ulong, [])):
        if not hasattr(zb_, name):
            zb_._f(name, res, *args)
    return zb_


def _deflate_with_header(zz, d, level, hdr):
    s = refz.ZStream()
    assert zz.deflateInit2_(C.byref(s), level, 8, 31, 8, 0, zz.version, C.sizeof(refz.ZStream)) == 0
    assert zz.deflateSetHeader(C.byref(s), C.byref(hdr)) == 0
    src = C.create_string_buffer(d, len(d))
    dst = C.create_string_buffer(len(d) + 4096)
    s.next_in, s.avail_in, s.next_out, s.avail_out = C.addressof(src), len(d), C.addressof(dst), len(dst)
    assert zz.deflate(C.byref(s), refz.Z_FINISH) == refz.Z_STREAM_END
    out = dst.raw[:len(dst) - s.avail_out]
    zz.deflateEnd(C.byref(s))
    return out


def test_advanced_stream_functions(z):
    """deflateSetHeader / inflateGetHeader / deflatePending / deflateCopy / inflateCopy / inflateGetDictionary /
    inflateValidate / zlibCompileFlags next to the reference (zlib.h "advanced functions")."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    z = _advanced(z)
    ref = _advanced(refz.ref())
    assert z.zlibCompileFlags() & 0xff == ref.zlibCompileFlags() & 0xff           # the type-size fields
    d = refz.gen(200000, refz.GEN_TEXT, seed=91)
    extra, name, comment = C.create_string_buffer(b"\x41\x42\x03\x00xyz", 7), C.create_string_buffer(b"file.txt"), C.create_string_buffer(b"made on a GPU")
    for hcrc in (0, 1):
        hdr = GzHeader(1, 1234567890, 0, 7, C.addressof(extra), 7, 0, C.addressof(name), 0, C.addressof(comment), 0, hcrc, 0)
        a, b = _deflate_with_header(z, d, 6, hdr), _deflate_with_header(ref, d, 6, hdr)
        assert a == b, (hcrc, a[:40].hex(), b[:40].hex())
        for zz in (z, ref):                                                        # ... and read back
            s = refz.ZStream()
            assert zz.inflateInit2_(C.byref(s), 31, zz.version, C.sizeof(refz.ZStream)) == 0
            eb, nb, cb = C.create_string_buffer(16), C.create_string_buffer(32), C.create_string_buffer(6)
            got = GzHeader(0, 0, 0, 0, C.addressof(eb), 0, 16, C.addressof(nb), 32, C.addressof(cb), 6, 0, 0)
            assert zz.inflateGetHeader(C.byref(s), C.byref(got)) == 0
            src, dst = C.create_string_buffer(a, len(a)), C.create_string_buffer(len(d) + 64)
            s.next_in, s.avail_in, s.next_out, s.avail_out = C.addressof(src), len(a), C.addressof(dst), len(dst)
            assert zz.inflate(C.byref(s), refz.Z_NO_FLUSH) == refz.Z_STREAM_END and dst.raw[:len(d)] == d   # (Z_FINISH: the reference keeps no window)
            assert (got.done, got.text, got.time, got.os, got.extra_len, got.hcrc) == (1, 1, 1234567890, 7, 7, hcrc)
            assert eb.raw[:7] == extra.raw[:7] and nb.value == b"file.txt" and cb.raw == b"made o"   # comm_max reached: not terminated
            # the window after the stream: its last 32 KiB
            win, wl = C.create_string_buffer(32768), C.c_uint(0)
            assert zz.inflateGetDictionary(C.byref(s), win, C.byref(wl)) == 0
            if zz is z:                                       # (the reference keeps no window when one call decodes the whole stream)
                assert wl.value == 32768 and win.raw == d[-32768:]
            zz.inflateEnd(C.byref(s))
    # deflatePending, deflateCopy: a copy taken mid-stream finishes to the same bytes
    s = refz.ZStream()
    assert z.deflateInit2_(C.byref(s), 6, 8, 15, 8, 0, z.version, C.sizeof(refz.ZStream)) == 0
    src, dst, dst2 = C.create_string_buffer(d, len(d)), C.create_string_buffer(len(d) + 4096), C.create_string_buffer(len(d) + 4096)
    s.next_in, s.avail_in, s.next_out, s.avail_out = C.addressof(src), 100000, C.addressof(dst), 10
    assert z.deflate(C.byref(s), refz.Z_SYNC_FLUSH) == 0 and s.avail_out == 0
    pend, bits = C.c_uint(0), C.c_int(9)
    assert z.deflatePending(C.byref(s), C.byref(pend), C.byref(bits)) == 0 and pend.value > 0 and bits.value == 0
    s2 = refz.ZStream()
    assert z.deflateCopy(C.byref(s2), C.byref(s)) == 0
    outs = []
    for st, buf in ((s, dst), (s2, dst2)):
        st.next_in, st.avail_in, st.next_out, st.avail_out = C.addressof(src) + 100000, len(d) - 100000, C.addressof(buf) + 10, len(buf) - 10
        assert z.deflate(C.byref(st), refz.Z_FINISH) == refz.Z_STREAM_END
        outs.append(dst.raw[:10] + buf.raw[10:len(buf) - st.avail_out])
        z.deflateEnd(C.byref(st))
    assert outs[0] == outs[1] and ref.inflate_all(outs[0], refz.WRAP_ZLIB, cap=len(d) + 64)[2] == d
    # inflateCopy mid-stream; inflateValidate(0) lets a damaged trailer pass (inflate.c:1495)
    stream = ref.deflate_stream(d, 6, 0, refz.WRAP_ZLIB, 50000)
    s = refz.ZStream()
    assert z.inflateInit2_(C.byref(s), 15, z.version, C.sizeof(refz.ZStream)) == 0
    src, o1, o2 = C.create_string_buffer(stream, len(stream)), C.create_string_buffer(len(d) + 64), C.create_string_buffer(len(d) + 64)
    s.next_in, s.avail_in, s.next_out, s.avail_out = C.addressof(src), len(stream) // 2, C.addressof(o1), len(o1)
    assert z.inflate(C.byref(s), 0) == 0
    got1 = len(o1) - s.avail_out
    s2 = refz.ZStream()
    assert z.inflateCopy(C.byref(s2), C.byref(s)) == 0
    C.memmove(o2, o1, got1)
    for st, buf in ((s, o1), (s2, o2)):
        st.next_in, st.avail_in, st.next_out, st.avail_out = C.addressof(src) + len(stream) // 2, len(stream) - len(stream) // 2, C.addressof(buf) + got1, len(buf) - got1
        assert z.inflate(C.byref(st), refz.Z_FINISH) == refz.Z_STREAM_END and buf.raw[:len(d)] == d
        z.inflateEnd(C.byref(st))
    bad = bytearray(stream); bad[-1] ^= 0xff
    for check, want in ((1, refz.Z_DATA_ERROR), (0, refz.Z_STREAM_END)):
        s = refz.ZStream()
        assert z.inflateInit2_(C.byref(s), 15, z.version, C.sizeof(refz.ZStream)) == 0 and z.inflateValidate(C.byref(s), check) == 0
        src, o1 = C.create_string_buffer(bytes(bad), len(bad)), C.create_string_buffer(len(d) + 64)
        s.next_in, s.avail_in, s.next_out, s.avail_out = C.addressof(src), len(bad), C.addressof(o1), len(o1)
        assert z.inflate(C.byref(s), refz.Z_FINISH) == want and z.inflateMark(C.byref(s)) == -65536
        z.inflateEnd(C.byref(s))


def test_reference_gun_linked_against_product(tmp_path):
    """The reference's examples/gun.c — gunzip through inflateBack() (infback.c:242) and crc32(), several members
    per file — compiled against the reference's zlib.h and linked to libzb200.so."""
    import gzip
    exe = os.path.join(refz.ROOT, "tests", "_bin", "gun_b200")
    if not os.path.exists(exe):
        pytest.skip("gun_b200 was not prebuilt (needs /root/reference at build time)")
    parts = [refz.gen(n, refz.GEN_TEXT, seed=n) for n in (300000, 0, 70000)]
    blob = b"".join(gzip.compress(p, 6) for p in parts)
    back = subprocess.run([exe], input=blob, capture_output=True, timeout=300)
    assert back.returncode == 0 and back.stdout == b"".join(parts), (back.returncode, back.stderr[:300])
    bad = bytearray(blob); bad[len(blob) // 3] ^= 0x20
    r = subprocess.run([exe], input=bytes(bad), capture_output=True, timeout=300)
    assert b"gun data error" in r.stderr                             # gun.c:595-630 reports, the exit code stays 0 for stdin
    r = subprocess.run([exe, "-t"], input=blob[:len(blob) // 2], capture_output=True, timeout=300)
    assert b"gun" in r.stderr and r.stdout == b""


@pytest.mark.parametrize("wrap", [refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP])
def test_inflate_big_slices_run_parallel(z, wrap):
    """inflate() fed in large slices: the runs between flush points that are complete in the input so far are decoded
    in one batch, the stream is re-based behind them, the wrapper's trailer is checked over the pieces (combine)."""
    ref = refz.ref() if refz.have_ref() else refz.oracle()
    d = refz.gen(12000000, refz.GEN_MARKOV, seed=71)
    s = ref.deflate_stream(d, 6, 0, wrap, 200000)                 # 60 full-flush runs
    for in_slice, out_slice in ((1000000, None), (700001, 3000000), (300000, 65536), (len(s) - 3, None), (len(s) - 9, 1 << 20)):
        ret, m, out, tin = z.inflate_all(s + b"after", wrap, cap=len(d) + 64, in_slice=in_slice, out_slice=out_slice)
        assert ret == refz.Z_STREAM_END and out == d and tin == len(s), (wrap, in_slice, out_slice, ret, m, len(out), tin, len(s))
    # damage: the reference's verdict, whatever the slicing
    for where, flip in ((len(s) // 2, 0x08), (len(s) - 2, 0x01), (40, 0x80)):
        bad = bytearray(s); bad[where] ^= flip
        want = ref.inflate_all(bytes(bad), wrap, cap=len(d) + 64)
        for in_slice in (1000000, 5000):
            got = z.inflate_all(bytes(bad), wrap, cap=len(d) + 64, in_slice=in_slice)
            assert got[0] == want[0] and (got[1] == want[1] or wrap == refz.WRAP_RAW), (wrap, where, in_slice, got[:2], want[:2])
            k = min(len(got[2]), len(want[2]))                  # whatever both delivered before stopping is the same bytes
            assert got[2][:k] == want[2][:k]
    # cut short: everything complete is delivered, Z_BUF_ERROR at the end (inflate.c:1259-1261)
    ret, m, out, tin = z.inflate_all(s[:len(s) * 2 // 3], wrap, cap=len(d) + 64, in_slice=900000)
    assert ret in (refz.Z_OK, refz.Z_BUF_ERROR) and d.startswith(out) and len(out) > len(d) // 2   # (Z_OK: the input-less last call still delivered)
    want = ref.inflate_all(s[:len(s) * 2 // 3], wrap, cap=len(d) + 64)[2]
    assert len(out) >= len(want) - 70000                        # (the reference also delivers the started block; this path stops at block boundaries)
    # sync flushes (runs need their predecessors) and a stream with no flush point at all stay correct
    if refz.have_ref():
        s2 = refz.ref().deflate_stream(d[:4000000], 6, 0, wrap, 300000, chunk_flush=[refz.Z_SYNC_FLUSH, refz.Z_FULL_FLUSH, refz.Z_SYNC_FLUSH])
        ret, m, out, tin = z.inflate_all(s2, wrap, cap=4000064, in_slice=500000)
        assert ret == refz.Z_STREAM_END and out == d[:4000000] and tin == len(s2)
    s3 = ref.deflate_stream(d[:3000000], 6, 0, wrap, 0)
    ret, m, out, tin = z.inflate_all(s3, wrap, cap=3000064, in_slice=400000)
    assert ret == refz.Z_STREAM_END and out == d[:3000000] and tin == len(s3)


@pytest.mark.parametrize("wrap", [refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP])
def test_inflate_big_slices_block_parallel(z, wrap):
    """inflate() fed a stream WITHOUT flush points in large slices: the chunks between dynamic block headers that are
    complete in the input so far are decoded in one batch (csrc/zb_inflate_blocks.cuh), the stream is re-based behind
    them at a block boundary INSIDE a byte, and so on; output, totals and verdicts are the reference's."""
    import zlib
    d = refz.gen(20000000, refz.GEN_MARKOV, seed=73) + refz.gen(4000000, refz.GEN_MIXED, seed=74)
    co = zlib.compressobj(6, zlib.DEFLATED, {refz.WRAP_RAW: -15, refz.WRAP_ZLIB: 15, refz.WRAP_GZIP: 31}[wrap])
    s = co.compress(d) + co.flush()
    ref = refz.ref() if refz.have_ref() else refz.oracle()
    for in_slice, out_slice in ((None, None), (2500000, None), (700001, 3000000), (300000, 1 << 20), (len(s) - 5, None)):
        ret, m, out, tin = z.inflate_all(s + b"after", wrap, cap=len(d) + 64, in_slice=in_slice, out_slice=out_slice)
        assert ret == refz.Z_STREAM_END and out == d and tin == len(s), (wrap, in_slice, out_slice, ret, m, len(out), tin, len(s))
    for where, flip in ((len(s) // 2, 0x08), (len(s) - 2, 0x01), (len(s) // 7, 0x80)):
        bad = bytearray(s); bad[where] ^= flip
        want = ref.inflate_all(bytes(bad), wrap, cap=len(d) + 64)
        for in_slice in (None, 1500000):
            got = z.inflate_all(bytes(bad), wrap, cap=len(d) + 64, in_slice=in_slice)
            assert got[0] == want[0] and (got[1] == want[1] or wrap == refz.WRAP_RAW), (wrap, where, in_slice, got[:2], want[:2])
            k = min(len(got[2]), len(want[2]))
            assert got[2][:k] == want[2][:k]
    ret, m, out, tin = z.inflate_all(s[:len(s) * 2 // 3], wrap, cap=len(d) + 64, in_slice=2000000)
    assert ret in (refz.Z_OK, refz.Z_BUF_ERROR) and d.startswith(out) and len(out) > len(d) // 2


def test_deflate_tune(z):
    """deflateTune (deflate.c:805-816): the four search parameters replace the level's table values — the
    reference's bytes for the same tuning at the lazy levels."""
    if not refz.have_ref():
        pytest.skip("oracle/_ref/libzref.so not built")
    ref = refz.ref()
    for lib in (z, ref):
        if not hasattr(lib, "deflateTune"):
            lib._f("deflateTune", C.c_int, C.POINTER(refz.ZStream), C.c_int, C.c_int, C.c_int, C.c_int)
    d = refz.gen(600000, refz.GEN_MIXED, seed=13)
    outs = {}
    for level, tune in ((6, (4, 8, 32, 64)), (9, (8, 16, 64, 32)), (6, (32, 258, 258, 1024)), (6, None)):
        for name, lib in (("z", z), ("ref", ref)):
            s = refz.ZStream()
            assert lib.deflateInit2_(C.byref(s), level, 8, 15, 8, 0, lib.version, C.sizeof(refz.ZStream)) == 0
            if tune:
                assert lib.deflateTune(C.byref(s), *tune) == 0
            src, dst = C.create_string_buffer(d, len(d)), C.create_string_buffer(len(d) + 4096)
            produced = 0
            for off in range(0, len(d), 262144):
                k = min(262144, len(d) - off)
                s.next_in, s.avail_in = C.addressof(src) + off, k
                s.next_out, s.avail_out = C.addressof(dst) + produced, len(dst) - produced
                r = lib.deflate(C.byref(s), refz.Z_FINISH if off + k >= len(d) else refz.Z_FULL_FLUSH)
                produced = len(dst) - s.avail_out
            assert r == refz.Z_STREAM_END
            lib.deflateEnd(C.byref(s))
            outs[(name, level, tune)] = dst.raw[:produced]
        assert outs[("z", level, tune)] == outs[("ref", level, tune)], (level, tune)
    assert outs[("z", 6, (4, 8, 32, 64))] != outs[("z", 6, None)]


def test_many_host_threads_distinct_streams(z):
    """zlib.h is re-entrant: one thread per z_stream, no shared mutable state (FAQ:151-160).  Eight host threads drive
    their own deflate / inflate streams (sliced calls, different levels and wrappers, dictionaries, big slices) and
    checksums through the one process-wide engine context at once; every result is the single-threaded one."""
    import concurrent.futures as cf
    ref = refz.ref() if refz.have_ref() else refz.oracle()
    base = refz.gen(6000000, refz.GEN_MARKOV, seed=81)

    def job(k):
        d = base[k * 300000:k * 300000 + 1500000 + 77 * k]
        wrap = (refz.WRAP_RAW, refz.WRAP_ZLIB, refz.WRAP_GZIP)[k % 3]
        level = (1, 6, 9, 4)[k % 4]
        s = z.deflate_stream(d, level, 0, wrap, chunk=(0, 100000)[k % 2], in_slice=(None, 70000)[k % 2], out_slice=(None, 50000)[(k // 2) % 2])
        want = ref.deflate_stream(d, level, 0, wrap, 0 if k % 2 == 0 else 100000)   # (chunk 0: no flush inside, the reference's one run)
        ok = s == want if level >= 4 else len(s) <= 1.03 * len(want) + 16
        ret, m, out, tin = z.inflate_all(s, wrap, cap=len(d) + 64, in_slice=(None, 400000, 16384)[k % 3])
        ok = ok and ret == refz.Z_STREAM_END and out == d and tin == len(s)
        ok = ok and z.crc32_z(0, d, len(d)) == refz.oracle().crc32(d) and z.adler32_z(1, d, len(d)) == refz.oracle().adler32(d)
        return ok

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(job, range(16)))
    assert all(results), results


def test_zlib_calls_from_many_host_threads(z):
    """The zlib names are re-entrant (zlib.h; FAQ:151-160: one thread per z_stream): host threads calling at the same time are
    spread over a pool of engine contexts (zb_zlib_api.cu api_ctx) and every one of them gets its own results — one-shot
    calls, checksums and streams fed in slices, mixed, checked against an independent zlib (Python's).  One stream is also
    handed from thread to thread between calls."""
    import threading
    import zlib
    errors = []

    def worker(t):
        try:
            rng = random.Random(1000 + t)
            for it in range(10):
                n = rng.choice((1000, 70000, 300000, 1 << 20, 3 << 20))
                d = refz.gen(n, rng.choice((refz.GEN_TEXT, refz.GEN_MARKOV, refz.GEN_MIXED)), seed=t * 100 + it)
                level = rng.choice((1, 6, 9))
                cap = z.compressBound(n)
                dst, dl = C.create_string_buffer(cap), C.c_ulong(cap)
                assert z.compress2(dst, C.byref(dl), d, n, level) == 0
                assert zlib.decompress(dst.raw[:dl.value]) == d
                back, bl = C.create_string_buffer(n + 1), C.c_ulong(n + 1)
                assert z.uncompress(back, C.byref(bl), zlib.compress(d, 6), len(zlib.compress(d, 6))) == 0 and back.raw[:bl.value] == d
                assert z.crc32(0, d, n) == zlib.crc32(d) and z.adler32(1, d, n) == zlib.adler32(d)
                s = z.deflate_stream(d, level, 0, refz.WRAP_GZIP, 0, in_slice=65536, out_slice=50000)
                assert zlib.decompress(s, 31) == d
                ret, msg, out, tin = z.inflate_all(zlib.compress(d, 9), refz.WRAP_ZLIB, cap=n + 64, in_slice=30000, out_slice=70000)
                assert ret == refz.Z_STREAM_END and out == d
        except BaseException as ex:                          # noqa: BLE001 (reported by the main thread)
            errors.append((t, repr(ex)))

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors, errors
    # one z_stream, every call from another thread (and so, in turn, from another context of the pool)
    d = refz.gen(2 << 20, refz.GEN_MARKOV, seed=5)
    comp = zlib.compress(d, 6)
    strm = refz.ZStream()
    assert z.inflateInit2_(C.byref(strm), 15, z.version, C.sizeof(refz.ZStream)) == 0
    src, dst = C.create_string_buffer(comp, len(comp)), C.create_string_buffer(len(d) + 64)
    state = {"fed": 0, "produced": 0, "ret": 0}

    def one_call():
        k = min(100000, len(comp) - state["fed"])
        strm.next_in, strm.avail_in = C.addressof(src) + state["fed"], k
        state["fed"] += k
        while True:
            room = len(d) + 64 - state["produced"]
            strm.next_out, strm.avail_out = C.addressof(dst) + state["produced"], room
            state["ret"] = z.inflate(C.byref(strm), refz.Z_NO_FLUSH)
            state["produced"] += room - strm.avail_out
            if state["ret"] != 0 or strm.avail_in == 0:
                break

    while state["ret"] in (0, refz.Z_BUF_ERROR) and state["fed"] < len(comp):
        th = threading.Thread(target=one_call)
        th.start()
        th.join()
    assert state["ret"] == refz.Z_STREAM_END and dst.raw[:state["produced"]] == d
    z.inflateEnd(C.byref(strm))
