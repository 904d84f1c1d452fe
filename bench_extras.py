"""Side measurements for bench.py: the other legs of the hot path on bounded
workloads (device-resident, CUDA events on the launching stream), each with its
roofline fraction and the reference CPU timed beside it on rank 0.

  deflate_l1  config C4: level 1, Markov text, 256 KiB Z_FULL_FLUSH chunks
  deflate_l6  config C5: level 6, mixed-entropy data, same chunking (ratio vs reference)
  inflate     config C3: multi-member gzip, members log-uniform 64 KiB..1 MiB,
              compressed by the REFERENCE at level 6, one warp per member
"""
import concurrent.futures as cf
import ctypes as C
import json
import os
import time

CHUNK = 262144
SEED = 0x9E3779B97F4A7C15


def _peak():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def _time_steps(torch, stream, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def run(ctx, rank, world, nbytes, barrier, steps=3, warmup=1):
    import torch
    import torch.distributed as dist
    import refz
    import zlib_wasm_b200 as zb
    L = zb.lib()
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    threads = len(os.sched_getaffinity(0))
    ref = refz.ref() if refz.have_ref() else None
    peak = _peak()
    out = {}

    def maxr(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    blocks = nbytes // 65536
    # ---------------- deflate ----------------
    for name, level, kind in (("deflate_l1", 1, refz.GEN_MARKOV), ("deflate_l6", 6, refz.GEN_MIXED)):
        host = refz.gen(nbytes, kind, SEED, first_block=rank * blocks)
        d_in = torch.frombuffer(bytearray(host), dtype=torch.uint8).cuda()
        cap = L.zb200_deflate_bound(nbytes, CHUNK, zb.FRAME_RAW)
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")

        def step():
            r = L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), nbytes, CHUNK, level, 0, zb.FRAME_RAW, 1,
                                    d_out.data_ptr(), cap, None, d_tot.data_ptr(), sp)
            if r != 0:
                raise zb.ZB200Error(r, "zb200_deflate_dev")

        l0 = L.zb200_launch_count()
        ms = maxr(_time_steps(torch, stream, step, steps, warmup))
        launches = (L.zb200_launch_count() - l0) // (steps + warmup)
        clen = int(d_tot.item())
        e = {"value": round(world * nbytes / (ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms, 3),
             "bytes_per_gpu": nbytes, "chunk": CHUNK, "ratio": round(nbytes / clen, 4), "kernels_per_step": int(launches),
             "roofline": {"bound": "hbm", "achieved": round((nbytes + clen) / (ms * 1e-3) / 1e9, 2), "peak": peak,
                          "unit": "GB/s", "frac": round((nbytes + clen) / (ms * 1e-3) / 1e9 / peak, 5),
                          "algorithmic_bytes": nbytes + clen}}
        if rank == 0 and ref is not None:
            # reference on the same chunking: chunk-parallel over host threads, bounded sample
            nsamp = min(nbytes // CHUNK, 64 if level == 1 else 32)
            chunks = [host[i * CHUNK:(i + 1) * CHUNK] for i in range(nsamp)]
            t0 = time.perf_counter()
            with cf.ThreadPoolExecutor(max_workers=threads) as ex:
                sizes = list(ex.map(lambda c: len(ref.deflate_stream(c, level, 0, refz.WRAP_RAW, 0)), chunks))
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            ref.deflate_stream(chunks[0], level, 0, refz.WRAP_RAW, 0)
            dt1 = time.perf_counter() - t1
            # the GPU stream of the same prefix, for the ratio on identical chunking
            d_end = torch.zeros(nbytes // CHUNK, dtype=torch.int64, device="cuda")
            L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), nbytes, CHUNK, level, 0, zb.FRAME_RAW, 1,
                                d_out.data_ptr(), cap, d_end.data_ptr(), d_tot.data_ptr(), sp)
            torch.cuda.synchronize()
            ours = int(d_end[nsamp - 1].item())
            # the reference's per-chunk one-shot streams end in BFINAL instead of the 5-byte marker
            refsz = sum(sizes) + 3 * nsamp
            e["cpu_baseline"] = {"value": round(nsamp * CHUNK / dt / 1e9, 4), "unit": "GB/s", "cores": threads,
                                 "kind": "reference", "sample": "%d chunks of 256 KiB, chunk-parallel" % nsamp,
                                 "single_thread_value": round(CHUNK / dt1 / 1e9, 4)}
            e["size_vs_reference"] = round(ours / refsz, 5)
        out[name] = e
        del d_in, d_out

    # ---------------- inflate ----------------
    host = refz.gen(nbytes, refz.GEN_MARKOV, SEED ^ 0x33, first_block=rank * blocks)
    sizes, off, i = [], 0, 0
    zg = C.CDLL(os.path.join(refz.ROOT, "tools", "libzgen.so"))
    zg.zgen_member_size.restype = C.c_uint64
    zg.zgen_member_size.argtypes = [C.c_uint64] * 4
    while off < nbytes:
        k = min(int(zg.zgen_member_size(SEED, rank * 100000 + i, 65536, 1 << 20)), nbytes - off)
        sizes.append((off, k))
        off += k
        i += 1
    comp = ref if ref is not None else refz.oracle()
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        streams = list(ex.map(lambda s: comp.deflate_stream(host[s[0]:s[0] + s[1]], 6, 0, refz.WRAP_GZIP, 0), sizes))
    blob = b"".join(streams)
    members, coff = [], 0
    for (uoff, k), s in zip(sizes, streams):
        members.append(zb.Member(coff, len(s), uoff, k, 0, 0))
        coff += len(s)
    n_m = len(members)
    arr = (zb.Member * n_m)(*members)
    d_blob = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    d_members = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    d_res = torch.zeros(n_m * C.sizeof(zb.MemberResult), dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")

    def step_inf():
        r = L.zb200_inflate_dev(ctx.handle, d_blob.data_ptr(), d_out.data_ptr(), d_members.data_ptr(), n_m,
                                zb.WRAP_GZIP, 1, d_res.data_ptr(), sp)
        if r != 0:
            raise zb.ZB200Error(r, "zb200_inflate_dev")

    ms = maxr(_time_steps(torch, stream, step_inf, steps, warmup))
    res = (zb.MemberResult * n_m).from_buffer_copy(d_res.cpu().numpy().tobytes())
    ok = all(r.status == 0 for r in res) and bytes(d_out.cpu().numpy().tobytes()) == host
    e = {"value": round(world * nbytes / (ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms, 3),
         "bytes_per_gpu": nbytes, "members": n_m, "compressed_bytes": len(blob), "bit_exact": bool(ok),
         "roofline": {"bound": "hbm", "achieved": round((nbytes + len(blob)) / (ms * 1e-3) / 1e9, 2), "peak": peak,
                      "unit": "GB/s", "frac": round((nbytes + len(blob)) / (ms * 1e-3) / 1e9 / peak, 5),
                      "algorithmic_bytes": nbytes + len(blob)}}
    if rank == 0 and ref is not None:
        nsamp = min(n_m, 4 * threads)
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(max_workers=threads) as ex:
            outs = list(ex.map(lambda j: len(ref.inflate_all(streams[j], refz.WRAP_GZIP, cap=sizes[j][1] + 8)[2]), range(nsamp)))
        dt = time.perf_counter() - t0
        e["cpu_baseline"] = {"value": round(sum(outs) / dt / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": "reference",
                             "sample": "%d members, member-parallel" % nsamp}
    out["inflate"] = e
    return out
