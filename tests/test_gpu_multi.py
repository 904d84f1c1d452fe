"""GPU: the one-process multi-GPU entry points (zb200_multi_*) give the results of the single-GPU calls —
the same checksum, the same compressed bytes, the same member results — on however many GPUs the box shows."""
import ctypes as C

import pytest

import refz
import zlib_wasm_b200 as zb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def both():
    L = zb.lib()
    m = C.c_void_p()
    assert L.zb200_multi_create(None, 0, C.byref(m)) == 0, zb.last_error()
    ctx = zb.Context(0)
    yield L, m, ctx
    L.zb200_multi_destroy(m)
    ctx.close()


def test_multi_matches_single(both):
    L, m, ctx = both
    g = L.zb200_multi_count(m)
    if L.zb200_device_count() < 2:
        pytest.skip("one GPU visible: the multi-GPU entry points need two (bench.py --gpus N runs the same gate at N > 1)")
    assert g >= 2
    d = refz.gen(24 * 262144 + 12345, refz.GEN_MIXED, seed=77)
    o = refz.oracle()
    crc, adler = C.c_uint32(0), C.c_uint32(0)
    assert L.zb200_multi_checksum_host(m, d, len(d), 3, 0, 1, C.byref(crc), C.byref(adler)) == 0
    assert (crc.value, adler.value) == (o.crc32(d), o.adler32(d))
    assert L.zb200_multi_checksum_host(m, d, len(d), 3, 0x1234, 0x00050006, C.byref(crc), C.byref(adler)) == 0
    assert (crc.value, adler.value) == (o.crc32(d, 0x1234), o.adler32(d, 0x00050006))
    for level, frame, finish in ((6, zb.FRAME_ZLIB, 1), (1, zb.FRAME_GZIP, 1), (6, zb.FRAME_RAW, 0), (6, zb.FRAME_GZIP_MEMBERS, 1), (9, zb.FRAME_GZIP, 1)):
        for data in (d, d[:100], b"", d[:262144 * 3]):
            cap = L.zb200_deflate_bound(len(data), 262144, frame) + 4096
            out = C.create_string_buffer(cap)
            olen, a, c = C.c_size_t(cap), C.c_uint32(0), C.c_uint32(0)
            r = L.zb200_multi_deflate_host(m, data, len(data), 262144, level, 0, frame, finish, out, C.byref(olen), C.byref(a), C.byref(c))
            assert r == 0, zb.last_error()
            want = ctx.deflate_host(data, level, 0, frame, 262144, finish)
            assert out.raw[:olen.value] == want, (g, level, frame, finish, len(data), olen.value, len(want))
            assert (c.value, a.value) == (o.crc32(data), o.adler32(data))
    # carried history (zb200.h ZB200_CHUNK_CARRY): a GPU's first chunk is compressed behind the 32 KiB before its piece
    for level, frame, chunk, data in ((6, zb.FRAME_ZLIB, 262144, d), (1, zb.FRAME_RAW, 65536, d), (6, zb.FRAME_GZIP, 5000, d[:30011]), (6, zb.FRAME_RAW, 262144, d[:100])):
        cap = L.zb200_deflate_bound(len(data), chunk, frame) + 4096
        out = C.create_string_buffer(cap)
        olen = C.c_size_t(cap)
        r = L.zb200_multi_deflate_host(m, data, len(data), chunk, level, 0, frame | zb.CHUNK_CARRY, 1, out, C.byref(olen), None, None)
        assert r == 0, zb.last_error()
        want = ctx.deflate_host(data, level, 0, frame | zb.CHUNK_CARRY, chunk, 1)
        assert out.raw[:olen.value] == want, (g, level, frame, chunk, len(data), olen.value, len(want))
    # members: the file the GZIP_MEMBERS frame wrote, split by the discovery call, inflated on all GPUs
    cap = L.zb200_deflate_bound(len(d), 65536, zb.FRAME_GZIP_MEMBERS)
    out = C.create_string_buffer(cap)
    olen = C.c_size_t(cap)
    assert L.zb200_multi_deflate_host(m, d, len(d), 65536, 6, 0, zb.FRAME_GZIP_MEMBERS, 1, out, C.byref(olen), None, None) == 0
    blob = out.raw[:olen.value]
    nmax = len(d) // 65536 + 2
    tab = (zb.Member * nmax)()
    back = C.create_string_buffer(len(d) + 16)
    blen, nm, st = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
    assert L.zb200_gunzip_host(ctx.handle, blob, len(blob), back, len(d) + 16, C.byref(blen), C.byref(st), tab, nmax, C.byref(nm)) == 0
    assert st.value == 0 and back.raw[:blen.value] == d and nm.value == len(d) // 65536 + 1
    res = (zb.MemberResult * nm.value)()
    back2 = C.create_string_buffer(len(d) + 16)
    assert L.zb200_multi_inflate_host(m, blob, back2, tab, nm.value, zb.WRAP_GZIP, 1, res) == 0, zb.last_error()
    assert all(r.status == 0 for r in res) and back2.raw[:len(d)] == d
    assert sum(r.out_len for r in res) == len(d)


def test_zlib_api_over_all_devices(tmp_path):
    """$ZB200_DEVICES=all: compress2 / crc32_z / adler32_z of large buffers behind the zlib.h names run on every GPU and
    give what one GPU gives for the same chunking (read once per process: checked in child processes)."""
    import os
    import subprocess
    import sys
    if zb.lib().zb200_device_count() < 2:
        pytest.skip("one GPU visible: $ZB200_DEVICES=all then is the single-GPU path")
    code = r'''
import ctypes as C, sys, hashlib
sys.path.insert(0, %r); sys.path.insert(0, %r)
import refz, zlib_wasm_b200 as zb
z = refz.ZlibBinding(zb.LIB_PATH, "")
d = refz.gen(40 << 20, refz.GEN_MARKOV, seed=8)
cap = C.c_ulong(z.compressBound(len(d)))
dst = C.create_string_buffer(cap.value)
assert z.compress2(dst, C.byref(cap), d, len(d), 6) == 0
print(hashlib.sha256(dst.raw[:cap.value]).hexdigest(), z.crc32_z(0, d, len(d)), z.adler32_z(1, d, len(d)))
''' % (refz.ROOT, os.path.join(refz.ROOT, "tests"))
    outs = []
    # one GPU emits the reference's ONE run of blocks for a one-shot call (round 2); a run cannot be cut across GPUs, so the
    # sharded call emits the chunked stream — the bytes one GPU gives when it is told to chunk ($ZB200_SINGLE_RUN_MAX=0)
    for devices, single_run in ((None, "0"), ("all", None), (None, None)):
        env = dict(os.environ)
        env.pop("ZB200_DEVICES", None)
        env.pop("ZB200_SINGLE_RUN_MAX", None)
        if devices:
            env["ZB200_DEVICES"] = devices
        if single_run is not None:
            env["ZB200_SINGLE_RUN_MAX"] = single_run
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr[-400:]
        outs.append(r.stdout.strip().split())
    assert outs[0] == outs[1] and len(outs[0]) == 3
    assert outs[2][1:] == outs[0][1:] and outs[2][0] != outs[0][0]     # same checksums; the default is the one-run stream
