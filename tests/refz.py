"""ctypes bindings to the TEST-ONLY checkers: the unmodified reference compiled
into oracle/_ref/libzref.so (symbols carry zlib's own Z_PREFIX ``z_``), the
plain-C restatement oracle/liboracle.so, contrib/puff (second decoder) and the
synthetic-data generator tools/libzgen.so.

Nothing in here is imported by the product package.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

Z_NO_FLUSH, Z_PARTIAL_FLUSH, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_FINISH = 0, 1, 2, 3, 4
Z_OK, Z_STREAM_END, Z_NEED_DICT = 0, 1, 2
Z_ERRNO, Z_STREAM_ERROR, Z_DATA_ERROR, Z_MEM_ERROR, Z_BUF_ERROR, Z_VERSION_ERROR = -1, -2, -3, -4, -5, -6
Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED = 0, 1, 2, 3, 4
WRAP_RAW, WRAP_ZLIB, WRAP_GZIP = 0, 1, 2


class ZStream(C.Structure):
    """z_stream, zlib.h:90-110 (LP64 layout, 112 bytes)."""
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_uint), ("total_in", C.c_ulong),
                ("next_out", C.c_void_p), ("avail_out", C.c_uint), ("total_out", C.c_ulong),
                ("msg", C.c_char_p), ("state", C.c_void_p),
                ("zalloc", C.c_void_p), ("zfree", C.c_void_p), ("opaque", C.c_void_p),
                ("data_type", C.c_int), ("adler", C.c_ulong), ("reserved", C.c_ulong)]


def _wbits(wrap):
    return {WRAP_RAW: -15, WRAP_ZLIB: 15, WRAP_GZIP: 31, 3: 47}[wrap]


class ZlibBinding:
    """The zlib.h subset used by the tests, bound to any library exporting it
    under an optional symbol prefix (``z_`` for libzref, none for the product)."""

    def __init__(self, path, prefix=""):
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        self.prefix = prefix
        L = self
        f = L._f
        f("zlibVersion", C.c_char_p)
        f("deflateInit2_", C.c_int, C.POINTER(ZStream), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int)
        f("deflateInit_", C.c_int, C.POINTER(ZStream), C.c_int, C.c_char_p, C.c_int)
        f("deflate", C.c_int, C.POINTER(ZStream), C.c_int)
        f("deflateEnd", C.c_int, C.POINTER(ZStream))
        f("deflateReset", C.c_int, C.POINTER(ZStream))
        f("deflateBound", C.c_ulong, C.POINTER(ZStream), C.c_ulong)
        f("inflateInit2_", C.c_int, C.POINTER(ZStream), C.c_int, C.c_char_p, C.c_int)
        f("inflateInit_", C.c_int, C.POINTER(ZStream), C.c_char_p, C.c_int)
        f("inflate", C.c_int, C.POINTER(ZStream), C.c_int)
        f("inflateEnd", C.c_int, C.POINTER(ZStream))
        f("inflateReset", C.c_int, C.POINTER(ZStream))
        f("compress", C.c_int, C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.c_ulong)
        f("compress2", C.c_int, C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.c_ulong, C.c_int)
        f("compressBound", C.c_ulong, C.c_ulong)
        f("uncompress", C.c_int, C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.c_ulong)
        f("uncompress2", C.c_int, C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.POINTER(C.c_ulong))
        f("crc32", C.c_ulong, C.c_ulong, C.c_void_p, C.c_uint)
        f("crc32_z", C.c_ulong, C.c_ulong, C.c_void_p, C.c_size_t)
        f("crc32_combine", C.c_ulong, C.c_ulong, C.c_ulong, C.c_long)
        f("crc32_combine_gen", C.c_ulong, C.c_long)
        f("crc32_combine_op", C.c_ulong, C.c_ulong, C.c_ulong, C.c_ulong)
        f("adler32", C.c_ulong, C.c_ulong, C.c_void_p, C.c_uint)
        f("adler32_z", C.c_ulong, C.c_ulong, C.c_void_p, C.c_size_t)
        f("adler32_combine", C.c_ulong, C.c_ulong, C.c_ulong, C.c_long)
        f("zError", C.c_char_p, C.c_int)
        f("deflateSetDictionary", C.c_int, C.POINTER(ZStream), C.c_void_p, C.c_uint)
        f("inflateSetDictionary", C.c_int, C.POINTER(ZStream), C.c_void_p, C.c_uint)
        self.version = self.zlibVersion()

    def _f(self, name, res, *args):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = res
        fn.argtypes = list(args)
        setattr(self, name, fn)

    # ---- helpers -------------------------------------------------------
    def deflate_stream(self, data, level=6, strategy=0, wrap=WRAP_ZLIB, chunk=0, mem_level=8,
                       out_slice=None, in_slice=None, dictionary=None, chunk_flush=Z_FULL_FLUSH, last_flush=Z_FINISH):
        """deflate(Z_FULL_FLUSH) per `chunk` bytes, Z_FINISH on the last one
        (SURVEY.md appendix C.1).  chunk=0: one Z_FINISH call.  in_slice /
        out_slice feed avail_in / avail_out in small pieces (zpipe style)."""
        data = bytes(data)
        n = len(data)
        strm = ZStream()
        r = self.deflateInit2_(C.byref(strm), level, 8, _wbits(wrap), mem_level, strategy,
                               self.version, C.sizeof(ZStream))
        if r != Z_OK:
            raise RuntimeError("deflateInit2 %d" % r)
        if dictionary is not None:                         # deflate.c:550-632
            r = self.deflateSetDictionary(C.byref(strm), bytes(dictionary), len(dictionary))
            if r != Z_OK:
                self.deflateEnd(C.byref(strm))
                raise RuntimeError("deflateSetDictionary %d" % r)
        src = C.create_string_buffer(data, max(n, 1))
        cap = n + (n >> 8) + 1024 + 16 * (n // chunk + 1 if chunk else 1)
        dst = C.create_string_buffer(cap)
        base_in, base_out = C.addressof(src), C.addressof(dst)
        chunk = chunk or max(n, 1)
        off = 0
        produced = 0
        while True:
            k = min(chunk, n - off)
            flush = last_flush if off + k >= n else (chunk_flush[(off // chunk) % len(chunk_flush)] if isinstance(chunk_flush, (list, tuple)) else chunk_flush)
            fed = 0
            while True:
                step = k - fed if not in_slice else min(in_slice, k - fed)
                strm.next_in = base_in + off + fed
                strm.avail_in = step
                fed += step
                fl = flush if fed == k else Z_NO_FLUSH
                while True:
                    room = cap - produced if not out_slice else min(out_slice, cap - produced)
                    strm.next_out = base_out + produced
                    strm.avail_out = room
                    r = self.deflate(C.byref(strm), fl)
                    if r not in (Z_OK, Z_STREAM_END, Z_BUF_ERROR):
                        self.deflateEnd(C.byref(strm))
                        raise RuntimeError("deflate %d" % r)
                    produced += room - strm.avail_out
                    if strm.avail_out != 0 or r == Z_STREAM_END:
                        break
                assert strm.avail_in == 0
                if fed == k:
                    break
            off += k
            if flush == Z_FINISH:
                assert r == Z_STREAM_END, r
                break
            if off >= n:                                   # last_flush != Z_FINISH: the stream is left open at a flush point
                break
        adler = strm.adler
        self.deflateEnd(C.byref(strm))
        return dst.raw[:produced]

    def inflate_all(self, data, wrap=WRAP_ZLIB, cap=None, in_slice=None, out_slice=None, dictionary=None):
        """Returns (ret, msg, output bytes, total_in).  `dictionary` is handed over when inflate asks
        for it (Z_NEED_DICT), or before the first call for a raw stream (inflate.c:1278-1312)."""
        data = bytes(data)
        n = len(data)
        cap = cap if cap is not None else max(64, n * 1100 + 1024)
        strm = ZStream()
        r = self.inflateInit2_(C.byref(strm), _wbits(wrap), self.version, C.sizeof(ZStream))
        if r != Z_OK:
            raise RuntimeError("inflateInit2 %d" % r)
        src = C.create_string_buffer(data, max(n, 1))
        dst = C.create_string_buffer(max(cap, 1))
        base_in, base_out = C.addressof(src), C.addressof(dst)
        fed = produced = 0
        ret = Z_OK
        if dictionary is not None and wrap == WRAP_RAW:
            r = self.inflateSetDictionary(C.byref(strm), bytes(dictionary), len(dictionary))
            if r != Z_OK:
                self.inflateEnd(C.byref(strm))
                raise RuntimeError("inflateSetDictionary %d" % r)
            dictionary = None
        while True:
            step = n - fed if not in_slice else min(in_slice, n - fed)
            strm.next_in = base_in + fed
            strm.avail_in = step
            stalled = False
            while True:
                room = cap - produced if not out_slice else min(out_slice, cap - produced)
                strm.next_out = base_out + produced
                strm.avail_out = room
                ret = self.inflate(C.byref(strm), Z_NO_FLUSH)
                produced += room - strm.avail_out
                if ret == Z_BUF_ERROR and strm.avail_in == 0 and fed + step < n:
                    ret = Z_OK                              # nothing pending and no input left in this slice: not an error
                    break                                   # (zlib.h:522-527; examples/zpipe.c ignores it the same way)
                if ret == 2 and dictionary is not None:    # Z_NEED_DICT
                    ret = self.inflateSetDictionary(C.byref(strm), bytes(dictionary), len(dictionary))
                    dictionary = None
                    if ret == Z_OK:
                        continue
                if ret != Z_OK:
                    break
                if strm.avail_out != 0:
                    break
                if produced >= cap:
                    stalled = True
                    break
            fed += step - strm.avail_in
            if ret != Z_OK or stalled:
                break
            if fed >= n:
                # one more call with no input reports Z_BUF_ERROR for a truncated stream
                strm.avail_in = 0
                strm.next_out = base_out + produced
                strm.avail_out = cap - produced
                ret = self.inflate(C.byref(strm), Z_NO_FLUSH)
                produced += (cap - produced) - strm.avail_out
                break
        msg = strm.msg.decode() if strm.msg else ""
        tin = strm.total_in
        self.inflateEnd(C.byref(strm))
        return ret, msg, dst.raw[:produced], tin


_cache = {}


def _load(key, path, builder=None):
    if key not in _cache:
        if not os.path.exists(path):
            raise FileNotFoundError("%s missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        _cache[key] = builder(path) if builder else C.CDLL(path, mode=C.RTLD_LOCAL)
    return _cache[key]


def ref():
    """The unmodified reference (zlib 1.3.1.1-motley) — oracle/_ref/libzref.so."""
    return _load("ref", os.path.join(ROOT, "oracle", "_ref", "libzref.so"), lambda p: ZlibBinding(p, "z_"))


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libzref.so"))


class RpResult(C.Structure):
    _fields_ = [("best_s", C.c_double), ("first_s", C.c_double), ("out_bytes", C.c_uint64), ("threads", C.c_int32), ("err", C.c_int32)]


class RpMember(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("in_len", C.c_uint64), ("out_off", C.c_uint64), ("out_cap", C.c_uint64)]


def _bind_refpool(path):
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    L.rp_deflate.restype = i
    L.rp_deflate.argtypes = [vp, sz, sz, i, i, i, i, vp, sz, vp, C.POINTER(RpResult)]
    L.rp_inflate.restype = i
    L.rp_inflate.argtypes = [vp, vp, sz, i, i, i, vp, vp, C.POINTER(RpResult)]
    L.rp_inflate_open.restype = i
    L.rp_inflate_open.argtypes = [vp, vp, sz, i, i, i, vp, vp, C.POINTER(RpResult)]
    L.rp_deflate_members.restype = i
    L.rp_deflate_members.argtypes = [vp, vp, sz, i, i, i, i, i, vp, vp, C.POINTER(RpResult)]
    L.rp_checksum.restype = i
    L.rp_checksum.argtypes = [vp, sz, i, i, i, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(RpResult)]
    L.rp_compress2_whole.restype = i
    L.rp_compress2_whole.argtypes = [vp, sz, i, i, i, i, vp, sz, C.POINTER(sz), C.POINTER(C.c_double)]
    return L


def refpool():
    """oracle/_ref/librefpool.so: the reference behind a pthread pool (BASELINE.md §3 CPU baseline)."""
    return _load("refpool", os.path.join(ROOT, "oracle", "_ref", "librefpool.so"), _bind_refpool)


def have_refpool():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librefpool.so"))


class Oracle:
    """oracle/liboracle.so.  Raw C entry points are bound as ``c_<name>``
    (without the zo_ prefix); the methods below are convenience helpers."""

    def __init__(self, path):
        L = self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        u32, u64, sz, vp = C.c_uint32, C.c_uint64, C.c_size_t, C.c_void_p
        sig = {
            "zo_crc32": (u32, u32, vp, sz), "zo_adler32": (u32, u32, vp, sz),
            "zo_multmodp": (u32, u32, u32), "zo_x2nmodp": (u32, u64, C.c_uint),
            "zo_crc32_combine": (u32, u32, u32, u64), "zo_crc32_combine_gen": (u32, u64),
            "zo_crc32_combine_op": (u32, u32, u32, u32), "zo_adler32_combine": (u32, u32, u32, C.c_int64),
            "zo_inflate_msg": (C.c_char_p, C.c_int),
            "zo_inflate": (C.c_int, vp, sz, vp, sz, C.c_int, C.POINTER(sz), C.POINTER(sz)),
            "zo_deflate_chunk": (sz, vp, sz, C.c_int, C.c_int, C.c_int, vp, sz),
            "zo_deflate_stream": (sz, vp, sz, C.c_int, C.c_int, C.c_int, sz, vp, sz),
            "zo_compress_bound": (sz, sz),
        }
        for k, v in sig.items():
            fn = getattr(L, k)
            fn.restype, fn.argtypes = v[0], list(v[1:])
            setattr(self, "c_" + k[3:], fn)

    def crc32(self, data, crc=0):
        return self.c_crc32(crc, bytes(data), len(data))

    def adler32(self, data, adler=1):
        return self.c_adler32(adler, bytes(data), len(data))

    def deflate_stream(self, data, level=6, strategy=0, wrap=WRAP_ZLIB, chunk=0):
        data = bytes(data)
        n = len(data)
        cap = n + (n >> 8) + 1024 + 16 * (n // chunk + 1 if chunk else 1)
        dst = C.create_string_buffer(cap)
        r = self.c_deflate_stream(data, n, level, strategy, wrap, chunk, dst, cap)
        if r == C.c_size_t(-1).value:
            raise RuntimeError("oracle deflate failed")
        return dst.raw[:r]

    def inflate_all(self, data, wrap=WRAP_ZLIB, cap=None):
        """Returns (err class, msg, output, consumed)."""
        data = bytes(data)
        n = len(data)
        cap = cap if cap is not None else max(64, n * 1100 + 1024)
        dst = C.create_string_buffer(max(cap, 1))
        used, made = C.c_size_t(0), C.c_size_t(0)
        e = self.c_inflate(data, n, dst, cap, wrap, C.byref(used), C.byref(made))
        return e, self.c_inflate_msg(e).decode(), dst.raw[:made.value], used.value


def oracle():
    return _load("oracle", os.path.join(ROOT, "oracle", "liboracle.so"), Oracle)


def puff():
    """contrib/puff as an independent decoder: int puff(dest,&destlen,src,&srclen)."""
    L = _load("puff", os.path.join(ROOT, "oracle", "_ref", "puff.so"))
    L.puff.restype = C.c_int
    L.puff.argtypes = [C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.POINTER(C.c_ulong)]
    return L


# ---- synthetic data ------------------------------------------------------
GEN_TEXT, GEN_MARKOV, GEN_RANDOM, GEN_MIXED, GEN_BYTES = 0, 1, 2, 3, 4
SEED = 0x9E3779B97F4A7C15


def gen(n, kind=GEN_TEXT, seed=SEED, first_block=0):
    L = _load("zgen", os.path.join(ROOT, "tools", "libzgen.so"))
    L.zgen_fill.restype = None
    L.zgen_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64]
    buf = C.create_string_buffer(max(n, 1))
    L.zgen_fill(buf, n, kind, seed, first_block)
    return buf.raw[:n]
