"""Device-resident deflate through zb200_deflate_dev, timed with CUDA events on the launching stream (the bench's headline
region at a chosen size): python tools/dev_time.py MiB level [markov|mixed] [steps]  — GB/s; run with ZB200_DUAL_STREAM=0 / 1 to see
what two sub-batches in flight give."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench_legs as BL  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

mib, level = int(sys.argv[1]), int(sys.argv[2])
gen = sys.argv[3] if len(sys.argv) > 3 else ("markov" if level < 4 else "mixed")
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
n = mib << 20
L = zb.lib()
ctx = zb.Context(0)
host = BL.host_alloc(L, n)
BL.fill(host, n, gen, 0)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_in.copy_(torch.frombuffer((C.c_uint8 * n).from_address(host), dtype=torch.uint8))
cap = L.zb200_deflate_bound(n, BL.CHUNK, zb.FRAME_RAW)
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
stream = torch.cuda.Stream()
sp = C.c_void_p(stream.cuda_stream)


def step():
    r = L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), n, BL.CHUNK, level, 0, zb.FRAME_RAW, 1, d_out.data_ptr(), cap, None, d_tot.data_ptr(), sp)
    assert r == 0, zb.last_error()


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(steps):
    step()
b.record(stream)
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
import hashlib
clen = int(d_tot.item())
digest = hashlib.sha1(d_out[:clen].cpu().numpy().tobytes()).hexdigest()[:16]
print("deflate_dev L%d %s %d MiB (ZB200_DUAL_STREAM=%s): %.2f ms  %.2f GB/s  %d bytes out  sha1 %s" % (
    level, gen, mib, os.environ.get("ZB200_DUAL_STREAM", "default"), ms, n / ms / 1e6, clen, digest), flush=True)
