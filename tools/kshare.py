"""Per-kernel time of each deflate / inflate leg on a bounded device-resident input, from the library's own CUDA events
(zb200_profile_*): python tools/kshare.py [MiB] [l1,l6,l9,l6t,inf]   (l6t = level 6 on markov text)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench_legs as BL  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
legs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["l1", "l6", "l6t", "inf"]
n = mib << 20
if os.environ.get("ZB_LIB"):
    zb.LIB_PATH = os.path.join(ROOT, os.environ["ZB_LIB"])   # a variant build (tools/build_variant.sh)
L = zb.lib()
ctx = zb.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = C.c_void_p(stream.cuda_stream)
host = L.zb200_host_alloc(n)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
cap = L.zb200_deflate_bound(n, BL.CHUNK, zb.FRAME_GZIP_MEMBERS)
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
d_end = torch.zeros(n // BL.CHUNK + 1, dtype=torch.int64, device="cuda")
view = torch.frombuffer((C.c_uint8 * n).from_address(host), dtype=torch.uint8)


def timed(fn, reps=3):
    fn()
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    ctx.profile(True)
    ctx.profile_read()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    k = ctx.profile_read()
    ctx.profile(False)
    return ms, {nm: round(v[0] / reps, 3) for nm, v in sorted(k.items(), key=lambda kv: -kv[1][0])}


cur = None
for leg in legs:
    gen = "markov" if leg in ("l1", "l2", "l3", "l6t", "inf") else "mixed"
    if gen != cur:
        BL.fill(host, n, gen, 0)
        d_in.copy_(view)
        torch.cuda.synchronize()
        cur = gen
    if leg == "inf":
        r = L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), n, BL.CHUNK, 6, 0, zb.FRAME_GZIP_MEMBERS, 1, d_out.data_ptr(), cap,
                                d_end.data_ptr(), d_tot.data_ptr(), sp)
        assert r == 0, zb.last_error()
        torch.cuda.synchronize()
        ends = d_end.cpu().tolist()[:n // BL.CHUNK]
        members, prev = [], 0
        for i, e in enumerate(ends):
            members.append(zb.Member(prev, e - prev, i * BL.CHUNK, BL.CHUNK, 0, 0, 0))
            prev = e
        arr = (zb.Member * len(members))(*members)
        d_m = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
        d_r = torch.zeros(len(members) * C.sizeof(zb.MemberResult), dtype=torch.uint8, device="cuda")
        d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
        ms, k = timed(lambda: L.zb200_inflate_dev(ctx.handle, d_out.data_ptr(), d_back.data_ptr(), d_m.data_ptr(), len(members),
                                                  zb.WRAP_GZIP, 1, d_r.data_ptr(), sp))
        ok = bool(torch.equal(d_back, d_in))
        print("inflate %d MiB (%d members): %.2f ms  %.1f GB/s  ok=%s  %s" % (mib, len(members), ms, n / ms / 1e6, ok, k), flush=True)
        continue
    level = {"l1": 1, "l2": 2, "l6": 6, "l6t": 6, "l9": 9, "l3": 3, "l4": 4}[leg]
    ms, k = timed(lambda: L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), n, BL.CHUNK, level, 0, zb.FRAME_RAW, 1, d_out.data_ptr(), cap,
                                              None, d_tot.data_ptr(), sp), 2 if level == 9 else 3)
    print("deflate %s %s %d MiB: %.2f ms  %.2f GB/s  ratio %.4f  %s" % (leg, gen, mib, ms, n / ms / 1e6, n / int(d_tot.item()), k), flush=True)
