"""zlib_wasm_b200 — B200-native zlib hot path (deflate / inflate / CRC-32 / Adler-32).

The product is the shared library ``libzb200.so`` in this directory: hand-written
CUDA kernels for sm_100a behind two C surfaces

  * ``zb200_*``  — the engine's C ABI (include/zb200.h): batched chunk / member /
    segment operations on device or host buffers, and
  * the zlib.h API of the reference (``deflateInit2_``, ``deflate``, ``inflate``,
    ``compress2``, ``uncompress``, ``crc32``, ``adler32`` ...) plus the
    ``zlib_*`` exports of the reference's src/wasm_module.c, so that the library
    is a drop-in for the reference's libz on this path.

This Python package is only a ctypes loader used by tests and bench.py; it holds
no algorithmic code and there is NO CPU fallback: if the library is missing, or
no B200 is visible, calls fail loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzb200.so")

OK = 0
ERR_NO_DEVICE, ERR_CUDA, ERR_PARAM, ERR_NOMEM, ERR_OUTPUT = -101, -102, -103, -104, -105
CRC32, ADLER32 = 1, 2
FRAME_RAW, FRAME_ZLIB, FRAME_GZIP, FRAME_GZIP_MEMBERS = 0, 1, 2, 3
EXACT_FAST = 0x200           # OR-ed into a frame: levels 1-3 with the reference's own chains, its bytes (zb200.h)
CHUNK_CARRY = 0x100          # OR-ed into a frame: every chunk compressed behind the 32 KiB before it (zb200.h)
WRAP_RAW, WRAP_ZLIB, WRAP_GZIP, WRAP_AUTO = 0, 1, 2, 3


class Member(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("in_len", C.c_uint64), ("out_off", C.c_uint64), ("out_cap", C.c_uint64),
                ("resume_bit", C.c_uint64), ("resume_out", C.c_uint64), ("dict_len", C.c_uint64)]


class MemberResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("wrap_kind", C.c_uint32), ("check", C.c_uint32), ("isize", C.c_uint32),
                ("out_len", C.c_uint64), ("in_used", C.c_uint64), ("resume_bit", C.c_uint64), ("resume_out", C.c_uint64)]


class DeflateOpts(C.Structure):
    _fields_ = [("level", C.c_int32), ("strategy", C.c_int32), ("window_bits", C.c_int32), ("mem_level", C.c_int32),
                ("dict_len", C.c_uint32), ("first_bit", C.c_uint32)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("ms", C.c_double), ("launches", C.c_uint64)]


class ZB200Error(RuntimeError):
    def __init__(self, code, what):
        super().__init__("%s failed: %d (%s)" % (what, code, last_error()))
        self.code = code


_lib = None

# every symbol include/zb200.h declares: (restype, argtypes)
_vp, _sz, _u32, _u64, _i = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int
_p32, _p64, _psz = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)
ABI = {
    "zb200_device_count": (_i,),
    "zb200_create": (_i, _i, C.POINTER(_vp)),
    "zb200_destroy": (None, _vp),
    "zb200_ctx_device": (_i, _vp),
    "zb200_sync": (_i, _vp, _vp),
    "zb200_last_error": (C.c_char_p,),
    "zb200_version": (C.c_char_p,),
    "zb200_host_alloc": (_vp, _sz),
    "zb200_host_free": (None, _vp),
    "zb200_launch_count": (_u64,),
    "zb200_profile_enable": (_i, _vp, _i),
    "zb200_profile_read": (_i, _vp, _vp, _sz, _psz),
    "zb200_checksum_dev": (_i, _vp, _vp, _sz, _i, _u32, _u32, _vp, _vp),
    "zb200_checksum_dev_sync": (_i, _vp, _vp, _sz, _i, _u32, _u32, _p32, _p32, _vp),
    "zb200_checksum_segments_dev": (_i, _vp, _vp, _vp, _vp, _sz, _i, _vp, _vp, _vp),
    "zb200_checksum_host": (_i, _vp, _vp, _sz, _i, _u32, _u32, _p32, _p32),
    "zb200_crc32_combine": (_u32, _u32, _u32, _u64),
    "zb200_crc32_combine_gen": (_u32, _u64),
    "zb200_crc32_combine_op": (_u32, _u32, _u32, _u32),
    "zb200_adler32_combine": (_u32, _u32, _u32, C.c_int64),
    "zb200_deflate_bound": (_sz, _sz, _sz, _i),
    "zb200_deflate_scratch_bytes": (_sz, _sz, _sz),
    "zb200_deflate_dev": (_i, _vp, _vp, _sz, _sz, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _vp),
    "zb200_deflate_host": (_i, _vp, _vp, _sz, _sz, _i, _i, _i, _i, _vp, _psz, _p32, _p32),
    "zb200_inflate_msg": (C.c_char_p, _i),
    "zb200_inflate_dev": (_i, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp, _vp),
    "zb200_inflate_host": (_i, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp),
    "zb200_selftest_tables": (_i, _vp, _vp, _vp, _sz, _vp),
    "zb200_multi_create": (_i, _vp, _i, C.POINTER(_vp)),
    "zb200_multi_destroy": (None, _vp),
    "zb200_multi_count": (_i, _vp),
    "zb200_multi_checksum_host": (_i, _vp, _vp, _sz, _i, _u32, _u32, _p32, _p32),
    "zb200_multi_deflate_host": (_i, _vp, _vp, _sz, _sz, _i, _i, _i, _i, _vp, _psz, _p32, _p32),
    "zb200_multi_inflate_host": (_i, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp),
    "zb200_inflate_stream_host": (_i, _vp, _vp, _sz, _i, _vp, _sz, _vp),
    "zb200_gunzip_host": (_i, _vp, _vp, _sz, _vp, _sz, _psz, C.POINTER(_i), _vp, _sz, _psz),
    "zb200_deflate_host_opts": (_i, _vp, _vp, _sz, _sz, _vp, _i, _i, _vp, _psz, _p32, _p32, _p32),
    "zb200_deflate_host_dict": (_i, _vp, _vp, _sz, _sz, _i, _i, _i, _vp, _vp, _vp, _vp),
}


def lib():
    """The loaded product library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing — run `python -m zlib_wasm_b200.build` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        for name, sig in ABI.items():
            fn = getattr(L, name)          # AttributeError here == ABI drift, fail loudly
            fn.restype = sig[0]
            fn.argtypes = list(sig[1:])
        _lib = L
    return _lib


def last_error():
    e = lib().zb200_last_error()
    return e.decode() if e else ""


class Context:
    """One engine context (one GPU)."""

    def __init__(self, device=0):
        self._h = _vp()
        r = lib().zb200_create(device, C.byref(self._h))
        if r != OK:
            raise ZB200Error(r, "zb200_create(%d)" % device)
        self.device = device

    def close(self):
        if self._h:
            lib().zb200_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def profile(self, on):
        lib().zb200_profile_enable(self._h, 1 if on else 0)

    def profile_read(self):
        """{kernel name: (ms summed over launches, launches)} since the last read."""
        arr = (KernelTime * 64)()
        n = C.c_size_t(0)
        r = lib().zb200_profile_read(self._h, arr, 64, C.byref(n))
        if r != OK:
            raise ZB200Error(r, "zb200_profile_read")
        return {arr[i].name.decode(): (arr[i].ms, int(arr[i].launches)) for i in range(n.value)}

    # ---- host-buffer conveniences used by the tests ----------------------
    def checksum_host(self, data, which=CRC32 | ADLER32, crc=0, adler=1):
        data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
        c, a = C.c_uint32(crc), C.c_uint32(adler)
        buf = (C.c_char * max(len(data), 1)).from_buffer_copy(data if len(data) else b"\0")
        r = lib().zb200_checksum_host(self._h, buf, len(data), which, crc, adler, C.byref(c), C.byref(a))
        if r != OK:
            raise ZB200Error(r, "zb200_checksum_host")
        return c.value, a.value

    def deflate_host(self, data, level=6, strategy=0, frame=FRAME_ZLIB, chunk=262144, finish=1):
        data = bytes(data)
        n = len(data)
        cap = lib().zb200_deflate_bound(n, chunk, frame)
        out = C.create_string_buffer(cap)
        olen = C.c_size_t(cap)
        ad, cr = C.c_uint32(0), C.c_uint32(0)
        r = lib().zb200_deflate_host(self._h, data, n, chunk, level, strategy, frame, finish, out, C.byref(olen),
                                     C.byref(ad), C.byref(cr))
        if r != OK:
            raise ZB200Error(r, "zb200_deflate_host")
        return out.raw[:olen.value]

    def inflate_host(self, blob, members, wrap=WRAP_GZIP, verify=1, out_size=None, prefill=None):
        """members: list of (in_off, in_len, out_off, out_cap[, resume_bit, resume_out[, dict_len]]).  `prefill`:
        bytes placed at the start of the output buffer before the call (preset dictionaries, resumed output).
        Returns (output bytes, [MemberResult])."""
        n = len(members)
        arr = (Member * max(n, 1))(*[Member(*(tuple(m) + (0, 0, 0))[:7]) for m in members])
        res = (MemberResult * max(n, 1))()
        if out_size is None:
            out_size = max([m[2] + m[3] for m in members] + [1])
        out = C.create_string_buffer(out_size)
        if prefill:
            C.memmove(out, bytes(prefill), len(prefill))
        r = lib().zb200_inflate_host(self._h, bytes(blob), out, arr, n, wrap, verify, res)
        if r != OK:
            raise ZB200Error(r, "zb200_inflate_host")
        return out.raw, list(res)[:n]


_default = {}


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
