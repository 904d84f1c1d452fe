"""zb200_deflate_host end to end (pinned host buffers, pieces pipelined over three streams) next to the same bytes compressed
device-resident: python tools/e2e_time.py MiB level [markov|mixed] [chunk bytes: 0 = ONE run of blocks]   — GB/s of both and the library's per-kernel times of the
host call (ZB200_PIPE_PIECE_MIB changes the piece size)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench_legs as BL  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

mib, level = int(sys.argv[1]), int(sys.argv[2])
gen = sys.argv[3] if len(sys.argv) > 3 else ("markov" if level < 4 else "mixed")
n = mib << 20
CH = int(sys.argv[4]) if len(sys.argv) > 4 else BL.CHUNK
if CH == 0:
    CH = n
L = zb.lib()
ctx = zb.Context(0)
host = BL.host_alloc(L, n)
BL.fill(host, n, gen, 0)
cap = L.zb200_deflate_bound(n, CH, zb.FRAME_RAW)
h_out = BL.host_alloc(L, cap)
olen = C.c_size_t(cap)
ad, cr = C.c_uint32(0), C.c_uint32(0)


def call():
    olen.value = cap
    r = L.zb200_deflate_host(ctx.handle, C.c_void_p(host), n, CH, level, 0, zb.FRAME_RAW, 1, C.c_void_p(h_out), C.byref(olen), C.byref(ad), C.byref(cr))
    assert r == 0, zb.last_error()


# the host link on this box (a pinned 1 GiB copy each way): the ceiling of the end-to-end figure
_probe = torch.empty(min(n, 1 << 30), dtype=torch.uint8, device="cuda")
_hv = torch.frombuffer((C.c_uint8 * _probe.numel()).from_address(host), dtype=torch.uint8)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); _probe.copy_(_hv, non_blocking=True); torch.cuda.synchronize(); t_in = time.perf_counter() - t0
    t0 = time.perf_counter(); _hv.copy_(_probe, non_blocking=True); torch.cuda.synchronize(); t_out = time.perf_counter() - t0
print("link probe: %.1f GB/s in, %.1f GB/s out" % (_probe.numel() / t_in / 1e9, _probe.numel() / t_out / 1e9), flush=True)
del _probe
call()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
ctx.profile(True); ctx.profile_read(); call(); k = ctx.profile_read(); ctx.profile(False)
ksum = sum(v[0] for v in k.values())
print("deflate_host L%d %s %d MiB: %.1f ms  %.2f GB/s end to end; kernels %.1f ms (%.2f GB/s): %s" % (
    level, gen, mib, min(ts) * 1e3, n / min(ts) / 1e9, ksum, n / ksum / 1e6,
    ", ".join("%s %.1f x%d" % (nm.replace("dfl_", "").replace("_kernel", ""), v[0], v[1]) for nm, v in sorted(k.items(), key=lambda kv: -kv[1][0])[:7])), flush=True)
