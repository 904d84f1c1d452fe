"""GPU parity: CRC-32 / Adler-32 kernels through the C ABI against the golden
vectors, the oracle and (full size) the compiled reference."""
import ctypes as C
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


def test_golden_edge_sizes(ctx, golden):
    cs = golden["checksums"]
    big = refz.gen(cs["data_n"], cs["data_kind"])
    for c in cs["cases"]:
        d = big[c["off"]:c["off"] + c["n"]]
        assert ctx.checksum_host(d) == (c["crc32"], c["adler32"]), c
        assert ctx.checksum_host(d, crc=0xdeadbeef, adler=0x12345678 % (65521 << 16) | 5) == (c["crc32_seeded"], c["adler32_seeded"]), c
    assert ctx.checksum_host(b"") == (0, 1)
    assert ctx.checksum_host(b"", which=zb.CRC32, crc=77)[0] == 77


def test_random_sizes_vs_oracle(ctx):
    o = refz.oracle()
    rng = random.Random(5)
    blob = refz.gen(40 << 20, refz.GEN_BYTES, seed=3)
    sizes = [rng.randrange(1, 300000) for _ in range(20)] + [8 << 20, (8 << 20) + 1, (8 << 20) - 1, 33554431, 40 << 20]
    for n in sizes:
        off = rng.randrange(0, len(blob) - n + 1)
        d = blob[off:off + n]
        assert ctx.checksum_host(d) == (o.crc32(d), o.adler32(d)), n
    # running-checksum chaining across calls (zlib.h:1711-1768 usage pattern)
    c, a = 0, 1
    for k in range(0, len(blob), 7 << 20):
        c, a = ctx.checksum_host(blob[k:k + (7 << 20)], crc=c, adler=a)
    assert (c, a) == (o.crc32(blob), o.adler32(blob))


def test_segments_unaligned_device(ctx):
    """Per-segment kernel on device memory with arbitrary (unaligned) offsets."""
    import torch
    o = refz.oracle()
    rng = random.Random(9)
    blob = refz.gen(6 << 20, refz.GEN_MIXED, seed=8)
    d = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    segs = []
    for _ in range(300):
        n = rng.choice([0, 1, 5, 15, 16, 17, 100, 4096, rng.randrange(0, 400000)])
        off = rng.randrange(0, len(blob) - n + 1)
        segs.append((off, n))
    off_t = torch.tensor([s[0] for s in segs], dtype=torch.int64, device="cuda")
    len_t = torch.tensor([s[1] for s in segs], dtype=torch.int64, device="cuda")
    crc_t = torch.zeros(len(segs), dtype=torch.int32, device="cuda")
    ad_t = torch.zeros(len(segs), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    r = zb.lib().zb200_checksum_segments_dev(ctx.handle, d.data_ptr(), off_t.data_ptr(), len_t.data_ptr(), len(segs), 3,
                                              crc_t.data_ptr(), ad_t.data_ptr(), C.c_void_p(s.cuda_stream))
    assert r == 0, zb.last_error()
    torch.cuda.synchronize()
    crc = crc_t.cpu().numpy().astype("uint32")
    ad = ad_t.cpu().numpy().astype("uint32")
    for i, (off, n) in enumerate(segs):
        piece = blob[off:off + n]
        assert (int(crc[i]), int(ad[i])) == (o.crc32(piece), o.adler32(piece)), (i, off, n)


def test_full_size_4gib_vs_reference(ctx):
    """BASELINE config C2 at full size: 4 GiB (exceeds uInt, needs the _z entry
    points), bit-exact against the compiled reference, plus the size-independent
    combine property crc(A||B) == combine(crc(A), crc(B), |B|) on the GPU values."""
    import torch
    if not refz.have_ref():
        pytest.skip("compiled reference not available")
    r = refz.ref()
    n = 4 << 30
    L = zb.lib()
    zg = C.CDLL(refz.ROOT + "/tools/libzgen.so")
    zg.zgen_fill.restype = None
    zg.zgen_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64]
    h = L.zb200_host_alloc(n)
    assert h
    try:
        zg.zgen_fill(h, n, refz.GEN_BYTES, refz.SEED, 0)
        crc, adler = C.c_uint32(), C.c_uint32()
        assert L.zb200_checksum_host(ctx.handle, h, n, 3, 0, 1, C.byref(crc), C.byref(adler)) == 0, zb.last_error()
        # reference over the same bytes, 8 threads + its own combine
        import bench_legs
        _, rc, ra = bench_legs.Cpu().checksum(h, n, bench_legs.host_threads(), 1)
        assert (crc.value, adler.value) == (rc, ra)
        # device-resident single launch over the whole 4 GiB and the split property
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        d.copy_(torch.frombuffer((C.c_uint8 * n).from_address(h), dtype=torch.uint8))
        c2, a2 = C.c_uint32(), C.c_uint32()
        assert L.zb200_checksum_dev_sync(ctx.handle, d.data_ptr(), n, 3, 0, 1, C.byref(c2), C.byref(a2), None) == 0
        assert (c2.value, a2.value) == (rc, ra)
        k = (3 << 30) + 12345
        ca, aa, cb, ab = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        assert L.zb200_checksum_dev_sync(ctx.handle, d.data_ptr(), k, 3, 0, 1, C.byref(ca), C.byref(aa), None) == 0
        assert L.zb200_checksum_dev_sync(ctx.handle, d.data_ptr() + k, n - k, 3, 0, 1, C.byref(cb), C.byref(ab), None) == 0
        assert L.zb200_crc32_combine(ca.value, cb.value, n - k) == rc
        assert L.zb200_adler32_combine(aa.value, ab.value, n - k) == ra
    finally:
        L.zb200_host_free(h)
