"""One pass of each non-checksum leg (deflate L1, deflate L6, inflate) on a
bounded device-resident input: the command profiled under ncu for the per-kernel
launch list in profiles/.  Usage: python tools/prof_legs.py [MiB] [legs]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 128
legs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["l1", "l6", "inf"]
n = mib << 20
if os.environ.get("ZB_LIB"):
    zb.LIB_PATH = os.path.join(ROOT, os.environ["ZB_LIB"])   # a variant build (tools/build_variant.sh)
L = zb.lib()
ctx = zb.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = C.c_void_p(stream.cuda_stream)
CH = int(os.environ.get("ZB_CH", 262144))


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


host = refz.gen(n, int(os.environ.get("ZB_GEN", refz.GEN_MARKOV)))
d_in = torch.frombuffer(bytearray(host), dtype=torch.uint8).cuda()
cap = L.zb200_deflate_bound(n, CH, zb.FRAME_GZIP_MEMBERS)
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
d_end = torch.zeros(n // CH + 1, dtype=torch.int64, device="cuda")
for leg, level in (("l1", 1), ("l6", 6)):
    if leg in legs:
        dt = timed(lambda: L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), n, CH, level, 0, zb.FRAME_RAW, 1,
                                               d_out.data_ptr(), cap, None, d_tot.data_ptr(), sp))
        print("deflate L%d: %.2f ms  %.2f GB/s  ratio %.3f" % (level, dt * 1e3, n / dt / 1e9, n / int(d_tot.item())), flush=True)
if "inf" in legs:
    # members made on the GPU (level 6 gzip members are byte-identical to the reference's)
    r = L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), n, CH, 6, 0, zb.FRAME_GZIP_MEMBERS, 1, d_out.data_ptr(), cap,
                            d_end.data_ptr(), d_tot.data_ptr(), sp)
    assert r == 0, zb.last_error()
    torch.cuda.synchronize()
    ends = d_end.cpu().tolist()[:n // CH]
    members, prev = [], 0
    for i, e in enumerate(ends):
        members.append(zb.Member(prev, e - prev, i * CH, CH, 0, 0))
        prev = e
    arr = (zb.Member * len(members))(*members)
    d_m = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    d_res = torch.zeros(len(members) * C.sizeof(zb.MemberResult), dtype=torch.uint8, device="cuda")
    d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
    dt = timed(lambda: L.zb200_inflate_dev(ctx.handle, d_out.data_ptr(), d_back.data_ptr(), d_m.data_ptr(), len(members),
                                           zb.WRAP_GZIP, 1, d_res.data_ptr(), sp))
    ok = bool(torch.equal(d_back, d_in))
    print("inflate: %d members  %.2f ms  %.2f GB/s  bit_exact=%s" % (len(members), dt * 1e3, n / dt / 1e9, ok), flush=True)
if "c3" in legs:
    # config C3's shape: members of 64 KiB .. 1 MiB (equal byte share), level 6 gzip, order shuffled
    classes = [65536, 131072, 262144, 524288, 1048576]
    per_class = (n // len(classes)) // 1048576 * 1048576
    tot = per_class * len(classes)
    d_blob = torch.empty(sum(L.zb200_deflate_bound(per_class, sz, zb.FRAME_GZIP_MEMBERS) for sz in classes), dtype=torch.uint8, device="cuda")
    members, blob_base = [], 0
    for ci, sz in enumerate(classes):
        nm = per_class // sz
        d_e = torch.zeros(nm, dtype=torch.int64, device="cuda")
        bound = L.zb200_deflate_bound(per_class, sz, zb.FRAME_GZIP_MEMBERS)
        r = L.zb200_deflate_dev(ctx.handle, d_in.data_ptr() + ci * per_class, per_class, sz, 6, 0, zb.FRAME_GZIP_MEMBERS, 1,
                                d_blob.data_ptr() + blob_base, bound, d_e.data_ptr(), d_tot.data_ptr(), sp)
        assert r == 0, zb.last_error()
        torch.cuda.synchronize()
        prev = 0
        for i, e in enumerate(d_e.cpu().tolist()):
            members.append(zb.Member(blob_base + prev, e - prev, ci * per_class + i * sz, sz, 0, 0))
            prev = e
        blob_base += bound
    order = sorted(range(len(members)), key=lambda i: (i * 2654435761) & 0xffffffff)
    members = [members[i] for i in order]
    arr = (zb.Member * len(members))(*members)
    d_m = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    d_res = torch.zeros(len(members) * C.sizeof(zb.MemberResult), dtype=torch.uint8, device="cuda")
    d_back = torch.empty(tot, dtype=torch.uint8, device="cuda")
    dt = timed(lambda: L.zb200_inflate_dev(ctx.handle, d_blob.data_ptr(), d_back.data_ptr(), d_m.data_ptr(), len(members),
                                           zb.WRAP_GZIP, 1, d_res.data_ptr(), sp), reps=3)
    ok = bool(torch.equal(d_back, d_in[:tot]))
    print("inflate C3 shape: %d members  %.2f ms  %.2f GB/s  bit_exact=%s" % (len(members), dt * 1e3, tot / dt / 1e9, ok), flush=True)
