"""Side measurements for bench.py: the other legs of the hot path on bounded
workloads (device-resident, CUDA events on the launching stream), each with its
roofline fraction and the reference CPU timed beside it on rank 0.

  deflate_l1  config C4: level 1, Markov text, 256 KiB Z_FULL_FLUSH chunks
  deflate_l6  config C5: level 6, mixed-entropy data, same chunking (ratio vs reference)
  inflate     config C3: multi-member gzip, members of 64 KiB..1 MiB at level 6 (made on
              the GPU; byte-identical to the reference's gzip members, re-checked on a
              sample), one warp per member, 4x the deflate workload size
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
        # end to end through the host entry point: pinned host input -> pinned host output, copies inside the timed region
        h_in, h_out = L.zb200_host_alloc(nbytes), L.zb200_host_alloc(cap)
        if h_in and h_out:
            C.memmove(h_in, host, nbytes)
            olen = C.c_size_t(cap)

            def step_host():
                olen.value = cap
                r = L.zb200_deflate_host(ctx.handle, C.c_void_p(h_in), nbytes, CHUNK, level, 0, zb.FRAME_RAW, 1, C.c_void_p(h_out),
                                         C.byref(olen), None, None)
                if r != 0:
                    raise zb.ZB200Error(r, "zb200_deflate_host")

            step_host()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                step_host()
            dt = maxr((time.perf_counter() - t0) / steps)
            e["e2e"] = {"value": round(world * nbytes / dt / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": nbytes,
                        "d2h_bytes_per_step": int(olen.value), "api": "zb200_deflate_host on pinned host memory (pieces pipelined over three streams)"}
            # ... and back: that ONE raw stream (no member table, no index) decoded at its flush points in parallel
            h_back = L.zb200_host_alloc(nbytes + 64)
            if h_back:
                q = zb.MemberResult()
                clen = int(olen.value)

                def step_back():
                    r = L.zb200_inflate_stream_host(ctx.handle, C.c_void_p(h_out), clen, zb.WRAP_RAW, C.c_void_p(h_back), nbytes + 64, C.byref(q))
                    if r != 0:
                        raise zb.ZB200Error(r, "zb200_inflate_stream_host")

                step_back()
                t0 = time.perf_counter()
                for _ in range(steps):
                    step_back()
                dt = maxr((time.perf_counter() - t0) / steps)
                okb = q.status == 0 and q.out_len == nbytes and C.string_at(h_back, 1 << 20) == host[:1 << 20] and \
                    C.string_at(h_back + nbytes - (1 << 20), 1 << 20) == host[-(1 << 20):]
                e["inflate_back_e2e"] = {"value": round(world * nbytes / dt / 1e9, 3), "unit": "GB/s", "bit_exact_ends": bool(okb),
                                         "api": "zb200_inflate_stream_host: one stream, runs between flush points found and decoded in one batch"}
                L.zb200_host_free(C.c_void_p(h_back))
        if h_in:
            L.zb200_host_free(C.c_void_p(h_in))
        if h_out:
            L.zb200_host_free(C.c_void_p(h_out))
        out[name] = e
        del d_in, d_out

    # ---------------- inflate ----------------
    # Members of 64 KiB .. 1 MiB (five size classes, equal byte share), gzip-wrapped, level 6.
    # They are produced on the GPU (FRAME_GZIP_MEMBERS): at level 6 that output is byte-identical
    # to the reference's one-shot gzip stream, which rank 0 re-checks on a few members below.
    inf_bytes = 4 * nbytes
    classes = [65536, 131072, 262144, 524288, 1048576]
    per_class = (inf_bytes // len(classes)) // 1048576 * 1048576
    inf_bytes = per_class * len(classes)
    host = refz.gen(inf_bytes, refz.GEN_MARKOV, SEED ^ 0x33, first_block=rank * (inf_bytes // 65536))
    d_plain = torch.frombuffer(bytearray(host), dtype=torch.uint8).cuda()
    bounds = [L.zb200_deflate_bound(per_class, sz, zb.FRAME_GZIP_MEMBERS) for sz in classes]
    d_blob = torch.empty(sum(bounds), dtype=torch.uint8, device="cuda")
    d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    members, blob_base, comp_total = [], 0, 0
    for ci, sz in enumerate(classes):
        nm = per_class // sz
        d_end = torch.zeros(nm, dtype=torch.int64, device="cuda")
        r = L.zb200_deflate_dev(ctx.handle, d_plain.data_ptr() + ci * per_class, per_class, sz, 6, 0, zb.FRAME_GZIP_MEMBERS, 1,
                                d_blob.data_ptr() + blob_base, bounds[ci], d_end.data_ptr(), d_tot.data_ptr(), sp)
        if r != 0:
            raise zb.ZB200Error(r, "zb200_deflate_dev(members)")
        torch.cuda.synchronize()
        ends = d_end.cpu().tolist()
        prev = 0
        for i, e in enumerate(ends):
            members.append(zb.Member(blob_base + prev, e - prev, ci * per_class + i * sz, sz, 0, 0))
            prev = e
        comp_total += prev
        blob_base += bounds[ci]
    # interleave the size classes so that the work list is not sorted by size
    order = sorted(range(len(members)), key=lambda i: (i * 2654435761) & 0xffffffff)
    members = [members[i] for i in order]
    n_m = len(members)
    arr = (zb.Member * n_m)(*members)
    d_members = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    d_res = torch.zeros(n_m * C.sizeof(zb.MemberResult), dtype=torch.uint8, device="cuda")
    d_out = torch.empty(inf_bytes, dtype=torch.uint8, device="cuda")

    def step_inf():
        r = L.zb200_inflate_dev(ctx.handle, d_blob.data_ptr(), d_out.data_ptr(), d_members.data_ptr(), n_m,
                                zb.WRAP_GZIP, 1, d_res.data_ptr(), sp)
        if r != 0:
            raise zb.ZB200Error(r, "zb200_inflate_dev")

    ms = maxr(_time_steps(torch, stream, step_inf, steps, warmup))
    res = (zb.MemberResult * n_m).from_buffer_copy(d_res.cpu().numpy().tobytes())
    ok = all(r.status == 0 for r in res) and bool(torch.equal(d_out, d_plain))
    e = {"value": round(world * inf_bytes / (ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms, 3),
         "bytes_per_gpu": inf_bytes, "members": n_m, "member_sizes": classes, "compressed_bytes": comp_total, "bit_exact": bool(ok),
         "roofline": {"bound": "hbm", "achieved": round((inf_bytes + comp_total) / (ms * 1e-3) / 1e9, 2), "peak": peak,
                      "unit": "GB/s", "frac": round((inf_bytes + comp_total) / (ms * 1e-3) / 1e9 / peak, 5),
                      "algorithmic_bytes": inf_bytes + comp_total}}
    # end to end through the host entry point: a multi-member file with the sizes mixed (the shuffled order above
    # is the file order: member i's output follows member i-1's), pinned host input -> pinned host output
    h_in, h_out = L.zb200_host_alloc(comp_total), L.zb200_host_alloc(inf_bytes)
    if h_in and h_out:
        torch.cuda.synchronize()
        blob_np = d_blob.cpu().numpy()
        file_members, ipos, opos = [], 0, 0
        for m in members:
            C.memmove(h_in + ipos, blob_np[m.in_off:m.in_off + m.in_len].tobytes(), m.in_len)
            file_members.append(zb.Member(ipos, m.in_len, opos, m.out_cap, 0, 0))
            ipos += m.in_len
            opos += m.out_cap
        arr2 = (zb.Member * n_m)(*file_members)
        res2 = (zb.MemberResult * n_m)()

        def step_host():
            r = L.zb200_inflate_host(ctx.handle, C.c_void_p(h_in), C.c_void_p(h_out), arr2, n_m, zb.WRAP_GZIP, 1, res2)
            if r != 0:
                raise zb.ZB200Error(r, "zb200_inflate_host")

        step_host()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        dt = maxr((time.perf_counter() - t0) / steps)
        ok2 = all(r.status == 0 for r in res2) and all(
            C.string_at(h_out + fm.out_off, fm.out_cap) == host[m.out_off:m.out_off + m.out_cap]
            for fm, m in list(zip(file_members, members))[:: max(1, n_m // 16)])
        e["e2e"] = {"value": round(world * inf_bytes / dt / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": int(comp_total),
                    "d2h_bytes_per_step": inf_bytes, "bit_exact": bool(ok2),
                    "api": "zb200_inflate_host on pinned host memory (pieces pipelined over three streams)"}
        # the same file with NO member table: members discovered on the device (zb200_gunzip_host)
        olen, nm, st = C.c_size_t(0), C.c_size_t(0), C.c_int(0)

        def step_scan():
            r = L.zb200_gunzip_host(ctx.handle, C.c_void_p(h_in), comp_total, C.c_void_p(h_out), inf_bytes, C.byref(olen), C.byref(st),
                                    None, 0, C.byref(nm))
            if r != 0:
                raise zb.ZB200Error(r, "zb200_gunzip_host")

        step_scan()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_scan()
        dt = maxr((time.perf_counter() - t0) / steps)
        ok3 = st.value == 0 and nm.value == n_m and olen.value == inf_bytes and all(
            C.string_at(h_out + fm.out_off, fm.out_cap) == host[m.out_off:m.out_off + m.out_cap]
            for fm, m in list(zip(file_members, members))[:: max(1, n_m // 16)])
        e["e2e_no_index"] = {"value": round(world * inf_bytes / dt / 1e9, 3), "unit": "GB/s", "members_found": int(nm.value), "bit_exact": bool(ok3),
                             "api": "zb200_gunzip_host on pinned host memory: member starts discovered on the device, one batch"}
    if h_in:
        L.zb200_host_free(C.c_void_p(h_in))
    if h_out:
        L.zb200_host_free(C.c_void_p(h_out))
    if rank == 0 and ref is not None:
        # sample members back on the host: identity with the reference's deflate, and the reference's inflate timed
        nsamp = min(n_m, 4 * threads)
        blob_h = d_blob.cpu().numpy()
        samples = []
        for j in range(nsamp):
            m = members[j]
            samples.append((bytes(blob_h[m.in_off:m.in_off + m.in_len]), m.out_off, m.out_cap))
        ident = all(s == ref.deflate_stream(host[o:o + k], 6, 0, refz.WRAP_GZIP, 0) for s, o, k in samples[:4])
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(max_workers=threads) as ex:
            outs = list(ex.map(lambda t: len(ref.inflate_all(t[0], refz.WRAP_GZIP, cap=t[2] + 8)[2]), samples))
        dt = time.perf_counter() - t0
        e["members_identical_to_reference_deflate"] = bool(ident)
        e["cpu_baseline"] = {"value": round(sum(outs) / dt / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": "reference",
                             "sample": "%d members, member-parallel" % nsamp}
    out["inflate"] = e
    return out
