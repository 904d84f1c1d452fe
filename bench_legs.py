"""The legs of BASELINE.json's metric, each at its configuration's size, shared by both arms
of bench.py:

  C4  deflate level 1 on markov8g (8 GiB, 256 KiB Z_FULL_FLUSH chunks)            -> the headline
  C5  deflate level 6 / 9 x Z_DEFAULT_STRATEGY / Z_FILTERED on mixed2g (2 GiB)
  C3  inflate of gz8g (8 GiB of gzip members, log-uniform 64 KiB .. 1 MiB, level 6)
  C2  CRC-32 + Adler-32 over bytes4g (4 GiB)

Strong scaling: a configuration's bytes are split evenly over the ranks (chunk c -> rank floor(c*G/chunks)).
The GPU arm measures, per leg: device-resident throughput (CUDA events on the launching stream), the share of
every kernel (the library's own per-kernel events), the roofline on algorithmic bytes, the end-to-end figure
through the host-buffer C ABI, the parity gates of SURVEY.md §8d, and (N = 1, rank 0) the reference CPU beside it:
the unmodified reference behind a pthread pool (oracle/_ref/librefpool.so), >= 256 MiB prefix, best of 3.
"""
import ctypes as C
import json
import os
import time

GIB = 1 << 30
MIB = 1 << 20
CHUNK = 262144
SEED = 0x9E3779B97F4A7C15
ROOT = os.path.dirname(os.path.abspath(__file__))

# name -> (generator, total bytes of the configuration, level, strategy)
DEFLATE_LEGS = {
    "deflate_l1": ("markov", 8 * GIB, 1, 0),
    "deflate_l6": ("mixed", 2 * GIB, 6, 0),
    "deflate_l6_filtered": ("mixed", 2 * GIB, 6, 1),
    "deflate_l9": ("mixed", 2 * GIB, 9, 0),
    "deflate_l9_filtered": ("mixed", 2 * GIB, 9, 1),
}
INFLATE_TOTAL = 8 * GIB
CHECKSUM_TOTAL = 4 * GIB
CPU_PREFIX = 256 * MIB            # ">= 256 MiB prefix"
CPU_SINGLE = 32 * MIB             # the single-thread figure: one z_stream on the first 32 MiB


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


_numa = {}


def gpu_numa(dev):
    """NUMA node of CUDA device `dev` and that node's CPUs (sysfs), or (None, None) when the platform does not say."""
    if dev in _numa:
        return _numa[dev]
    node, cpus = None, None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[dev]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else dev
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                       # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node >= 0:
            with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
        else:
            node = None
    except Exception:
        node, cpus = None, None
    if not cpus:                                              # sysfs does not say (containers often show -1): ask NVML for the GPU's CPUs
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[dev]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else dev
            words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx), (os.cpu_count() + 63) // 64)
            got = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1} & os.sched_getaffinity(0)
            if got and got != os.sched_getaffinity(0):
                cpus = got
        except Exception:
            pass
    _numa[dev] = (node, cpus or None)
    return _numa[dev]


def host_alloc(L, n, dev=None):
    """zb200_host_alloc with the calling thread on the GPU's own NUMA node while the pages are pinned (they are placed
    where the thread that touches them first runs): a DMA from the other socket's memory crosses the inter-socket link."""
    if dev is None:
        dev = int(os.environ.get("LOCAL_RANK", "0"))
    node, cpus = gpu_numa(dev)
    if not cpus:
        return L.zb200_host_alloc(n)
    old = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
        return L.zb200_host_alloc(n)
    finally:
        os.sched_setaffinity(0, old)


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def scaled(total, scale):
    """A configuration's size under --scale (whole MiB, at least 64)."""
    return max(64 * MIB, int(total * scale) // MIB * MIB)


def shard(total, rank, world, align):
    units = (total + align - 1) // align
    per, extra = divmod(units, world)
    first = rank * per + min(rank, extra)
    count = per + (1 if rank < extra else 0)
    return min(first * align, total), min((first + count) * align, total)


_GEN = {"text": 0, "markov": 1, "random": 2, "mixed": 3, "bytes": 4}


def zgen():
    L = C.CDLL(os.path.join(ROOT, "tools", "libzgen.so"))
    L.zgen_fill.restype = None
    L.zgen_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64]
    L.zgen_member_size.restype = C.c_uint64
    L.zgen_member_size.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
    return L


def fill(ptr, n, kind, first_byte):
    """Bytes [first_byte, first_byte + n) of the seeded logical buffer of `kind` (block-parallel generator)."""
    assert first_byte % 65536 == 0
    zgen().zgen_fill(ptr, n, _GEN[kind], SEED, first_byte // 65536)


def member_sizes(total, seed):
    """C3's member schedule: sizes log-uniform in [64 KiB, 1 MiB] on a grid of 65 values 64 KiB * 2^(k/16)
    (whole KiB), drawn by the generator's hash; the last member is cut so that the sizes sum to `total`."""
    zg = zgen()
    grid = sorted(set(int(65536 * 2 ** (k / 16.0)) // 1024 * 1024 for k in range(65)))
    out, left, i = [], total, 0
    import bisect
    while left > 0:
        v = int(zg.zgen_member_size(seed, i, 65536, 1048576))
        g = grid[max(0, bisect.bisect_right(grid, v) - 1)]
        g = min(g, left)
        out.append(g)
        left -= g
        i += 1
    return out


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference behind a pthread pool (oracle/_ref/librefpool.so)
# ---------------------------------------------------------------------------------------------------
class Cpu:
    """kind 'reference': oracle/_ref (the unmodified reference compiled in place).  The oracle port
    (oracle/liboracle.so) is the fallback when the reference could not be compiled here."""

    def __init__(self):
        import refz
        self.refz = refz
        self.threads = host_threads()
        if refz.have_refpool() and refz.have_ref():
            self.rp = refz.refpool()
            self.ref = refz.ref()
            self.kind = "reference"
        else:
            self.rp = None
            self.ref = None
            self.kind = "port"
            self.o = refz.oracle()

    # -- deflate: returns (seconds best-of-reps, bytes out, stream bytes or None, chunk_end list or None)
    def deflate(self, addr, n, level, strategy, threads, reps, want_stream=False):
        refz = self.refz
        nch = (n + CHUNK - 1) // CHUNK
        if self.rp is not None:
            res = refz.RpResult()
            cap = n + (n >> 3) + 1024 * nch + 4096
            out = C.create_string_buffer(cap) if want_stream else None
            ends = (C.c_uint64 * nch)()
            e = self.rp.rp_deflate(addr, n, CHUNK, level, strategy, threads, reps, out, cap if out else 0, ends, C.byref(res))
            if e != 0:
                raise RuntimeError("reference deflate failed: %d" % e)
            return res.best_s, int(res.out_bytes), out, list(ends)
        data = C.string_at(addr, n)
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            s = self.o.deflate_stream(data, level, strategy, refz.WRAP_RAW, CHUNK)
            best = min(best, time.perf_counter() - t0)
        return best, len(s), C.create_string_buffer(s, len(s)) if want_stream else None, None

    def inflate(self, addr, members, wbits, threads, reps, out_addr, allow_open=False):
        """members: [(in_off, in_len, out_off, out_cap)]; returns (seconds, bytes out)."""
        refz = self.refz
        if self.rp is None:
            data_all = None
            best, tot = 1e30, 0
            for _ in range(reps):
                t0 = time.perf_counter()
                tot = 0
                for (io, il, oo, oc) in members:
                    err, msg, back, used = self.o.inflate_all(C.string_at(addr + io, il), {31: 2, 15: 1, -15: 0}[wbits], cap=oc + 8)
                    C.memmove(out_addr + oo, back, len(back))
                    tot += len(back)
                best = min(best, time.perf_counter() - t0)
            return best, tot
        arr = (refz.RpMember * len(members))(*[refz.RpMember(*m) for m in members])
        res = refz.RpResult()
        fn = self.rp.rp_inflate_open if allow_open else self.rp.rp_inflate
        e = fn(addr, arr, len(members), wbits, threads, reps, out_addr, None, C.byref(res))
        if e != 0:
            raise RuntimeError("reference inflate failed: %d" % e)
        return res.best_s, int(res.out_bytes)

    def deflate_members(self, addr, spans, level, wbits, threads):
        """spans: [(in_off, in_len)] -> list of the reference's streams, one per span (member-parallel)."""
        refz = self.refz
        if self.rp is None:
            return [self.o.deflate_stream(C.string_at(addr + o, k), level, 0, {31: 2, 15: 1, -15: 0}[wbits], 0) for o, k in spans]
        mem, off = [], 0
        for o, k in spans:
            cap = k + (k >> 3) + 1024
            mem.append(refz.RpMember(o, k, off, cap))
            off += cap
        arr = (refz.RpMember * len(mem))(*mem)
        out = C.create_string_buffer(off)
        lens = (C.c_uint64 * len(mem))()
        res = refz.RpResult()
        e = self.rp.rp_deflate_members(addr, arr, len(mem), level, 0, wbits, threads, 1, out, lens, C.byref(res))
        if e != 0:
            raise RuntimeError("reference deflate of members failed: %d" % e)
        return [C.string_at(C.addressof(out) + m.out_off, lens[i]) for i, m in enumerate(mem)]

    def checksum(self, addr, n, threads, reps):
        refz = self.refz
        if self.rp is None:
            t0 = time.perf_counter()
            c, a = self.o.c_crc32(0, addr, n), self.o.c_adler32(1, addr, n)
            return time.perf_counter() - t0, c, a
        c, a, res = C.c_uint32(0), C.c_uint32(0), refz.RpResult()
        self.rp.rp_checksum(addr, n, 3, threads, reps, C.byref(c), C.byref(a), C.byref(res))
        return res.best_s, c.value, a.value

    def combine(self, parts):
        """[(crc, adler, n)] in order -> (crc, adler) with the REFERENCE's combine functions."""
        crc, adler = parts[0][0], parts[0][1]
        for c, a, k in parts[1:]:
            if self.ref is not None:
                crc = self.ref.crc32_combine(crc, c, k)
                adler = self.ref.adler32_combine(adler, a, k)
            else:
                crc = self.o.c_crc32_combine(crc, c, k)
                adler = self.o.c_adler32_combine(adler, a, k)
        return crc, adler


def cpu_deflate_baseline(cpu, addr, n_avail, level, strategy, reps=3):
    """The CPU baseline object of one deflate leg (prefix of the rank's shard)."""
    n = min(n_avail, CPU_PREFIX)
    sec, outb, _, _ = cpu.deflate(addr, n, level, strategy, cpu.threads, reps)
    n1 = min(n_avail, CPU_SINGLE)
    sec1, _, _, _ = cpu.deflate(addr, n1, level, strategy, 1, 1)
    return {"value": round(n / sec / 1e9, 4), "unit": "GB/s", "cores": cpu.threads, "kind": cpu.kind,
            "sample": "first %d MiB of the leg's input, 256 KiB Z_FULL_FLUSH chunks, pthread pool (one z_stream per thread, "
                      "deflateReset per chunk), best of %d" % (n // MIB, reps),
            "single_thread_value": round(n1 / sec1 / 1e9, 4), "single_thread_sample": "first %d MiB, one z_stream, one pass" % (n1 // MIB),
            "ratio": round(n / outb, 4)}


# ---------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------
class Gpu:
    def __init__(self, torch, dist, zb, ctx, stream, rank, world):
        self.torch, self.dist, self.zb, self.ctx, self.stream = torch, dist, zb, ctx, stream
        self.rank, self.world = rank, world
        self.L = zb.lib()
        self.sp = C.c_void_p(stream.cuda_stream)
        self.peak, self.peak_src = peak()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, *xs):
        if self.world > 1:
            t = self.torch.tensor(list(xs), dtype=self.torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            xs = [float(v) for v in t.cpu()]
        return xs if len(xs) > 1 else xs[0]

    def sumr(self, *xs):
        if self.world > 1:
            t = self.torch.tensor(list(xs), dtype=self.torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            xs = [float(v) for v in t.cpu()]
        return xs if len(xs) > 1 else xs[0]

    def time_steps(self, fn, steps, warmup):
        """ms per step: K steps between two events on the launching stream, barrier + synchronize on both sides."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        for _ in range(steps):
            fn()
        b.record(self.stream)
        self.barrier()
        return a.elapsed_time(b) / steps

    def kernel_shares(self, fn, steps=2):
        """Per-kernel ms per step from the library's own events (zb200_profile_*), one extra profiled pass."""
        self.ctx.profile(True)
        self.ctx.profile_read()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        got = self.ctx.profile_read()
        self.ctx.profile(False)
        return {k: {"ms_per_step": round(ms / steps, 4), "launches_per_step": ln // steps} for k, (ms, ln) in
                sorted(got.items(), key=lambda kv: -kv[1][0])}

    def wall_steps(self, fn, steps):
        fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        return self.maxr(dt)

    def roofline(self, algo_bytes, ms, kernels, traffic=None, note=None):
        ach = algo_bytes / (ms * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": round(ach, 2), "peak": self.peak, "unit": "GB/s", "frac": round(ach / self.peak, 5),
             "traffic": traffic, "peak_source": self.peak_src, "algorithmic_bytes_per_step": int(algo_bytes)}
        if kernels:
            tot = sum(v["ms_per_step"] for v in kernels.values()) or 1.0
            dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
            r["kernel"] = dom
            r["kernel_ms_per_step"] = kernels[dom]["ms_per_step"]
            r["kernel_share_of_step"] = round(kernels[dom]["ms_per_step"] / tot, 3)
            r["kernel_achieved"] = round(algo_bytes / (kernels[dom]["ms_per_step"] * 1e-3) / 1e9, 2)
        if note:
            r["note"] = note
        return r


def host_equals_device(torch, ptr, n, d_tensor, piece=256 * MIB):
    """memcmp of host bytes [ptr, ptr+n) against a device tensor, on the device, piece by piece."""
    if n == 0:
        return True
    tmp = torch.empty(min(piece, n), dtype=torch.uint8, device="cuda")
    for off in range(0, n, piece):
        k = min(piece, n - off)
        tmp[:k].copy_(torch.frombuffer((C.c_uint8 * k).from_address(ptr + off), dtype=torch.uint8))
        if not bool(torch.equal(tmp[:k], d_tensor[off:off + k])):
            return False
    return True


def deflate_leg(g, cpu, name, total, steps, warmup, host=None, keep=False, traffic=None, clocks=None):
    """One deflate configuration on this rank's shard.  Returns (entry, kept buffers or None)."""
    torch, zb, L = g.torch, g.zb, g.L
    kind, _, level, strategy = DEFLATE_LEGS[name]
    lo, hi = shard(total, g.rank, g.world, CHUNK)
    n = hi - lo
    own_host = host is None
    if own_host:
        host = host_alloc(L, n)
        if not host:
            raise RuntimeError("pinned allocation of %d bytes failed" % n)
        fill(host, n, kind, lo)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    h_view = torch.frombuffer((C.c_uint8 * n).from_address(host), dtype=torch.uint8)
    d_in.copy_(h_view)
    cap = L.zb200_deflate_bound(n, CHUNK, zb.FRAME_RAW)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    nch = (n + CHUNK - 1) // CHUNK
    d_end = torch.zeros(nch, dtype=torch.int64, device="cuda")

    def step():
        r = L.zb200_deflate_dev(g.ctx.handle, d_in.data_ptr(), n, CHUNK, level, strategy, zb.FRAME_RAW, 1,
                                d_out.data_ptr(), cap, d_end.data_ptr(), d_tot.data_ptr(), g.sp)
        if r != 0:
            raise zb.ZB200Error(r, "zb200_deflate_dev")

    if clocks:
        for _ in range(warmup):
            step()
        g.barrier()
        clocks.start()
        l0 = L.zb200_launch_count()
        ms = g.time_steps(step, steps, 0)
        launches = L.zb200_launch_count() - l0
        clocks.stop()
    else:
        l0 = L.zb200_launch_count()
        ms = g.time_steps(step, steps, warmup)
        launches = (L.zb200_launch_count() - l0) * steps // (steps + warmup)
    ms = g.maxr(ms)
    clen = int(d_tot.item())
    kernels = g.kernel_shares(step)
    tot_u, tot_c = g.sumr(float(n), float(clen))
    e = {"config": "C4" if level == 1 else "C5", "value": round(tot_u / (ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms, 3),
         "steps": steps, "warmup": warmup, "total_bytes": int(tot_u), "bytes_per_gpu": n, "chunk": CHUNK, "level": level,
         "strategy": ["Z_DEFAULT_STRATEGY", "Z_FILTERED"][strategy], "generator": kind,
         "ratio": round(tot_u / tot_c, 4), "compressed_bytes": int(tot_c), "gpu_launches": int(launches),
         "kernels": kernels,
         "roofline": g.roofline(n + clen, ms, kernels, traffic,
                                "achieved/frac = (U + C) of this rank's shard / the whole step (all kernels); kernel_* = the dominant kernel alone")}

    # ---- end to end: pinned host input -> pinned host output through the C ABI ----
    h_out = host_alloc(L, cap)
    if not h_out:
        raise RuntimeError("pinned allocation of %d bytes failed" % cap)
    olen = C.c_size_t(cap)
    crc_in, adler_in = C.c_uint32(0), C.c_uint32(0)

    def step_host():
        olen.value = cap
        r = L.zb200_deflate_host(g.ctx.handle, C.c_void_p(host), n, CHUNK, level, strategy, zb.FRAME_RAW, 1, C.c_void_p(h_out),
                                 C.byref(olen), C.byref(adler_in), C.byref(crc_in))
        if r != 0:
            raise zb.ZB200Error(r, "zb200_deflate_host")

    e2e_steps = max(2, min(steps, 3))
    dt = g.wall_steps(step_host, e2e_steps)
    tot_d2h = g.sumr(float(olen.value))
    e["e2e"] = {"value": round(tot_u / dt / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": int(tot_u), "d2h_bytes_per_step": int(tot_d2h),
                "steps": e2e_steps, "api": "zb200_deflate_host, pinned host buffers, pieces pipelined over three streams"}

    # ---- parity gates ----
    gates = {}
    # (1) the host path and the device path emit the same bytes
    torch.cuda.synchronize()
    same = olen.value == clen and host_equals_device(torch, h_out, clen, d_out)
    gates["host_path_equals_device_path"] = bool(same)
    # (2) full-size round trip: the rank's whole stream back through the run-parallel decoder
    h_back = host_alloc(L, n + 64)
    if h_back:
        q = zb.MemberResult()

        def step_back():
            r = L.zb200_inflate_stream_host(g.ctx.handle, C.c_void_p(h_out), clen, zb.WRAP_RAW, C.c_void_p(h_back), n + 64, C.byref(q))
            if r != 0:
                raise zb.ZB200Error(r, "zb200_inflate_stream_host")

        dtb = g.wall_steps(step_back, 1)
        ok = q.status == 0 and q.out_len == n and host_equals_device(torch, h_back, n, d_in)
        gates["round_trip_full_size"] = bool(ok)
        gates["round_trip_crc_matches_input_crc"] = bool(q.check == crc_in.value)
        e["inflate_back_e2e"] = {"value": round(tot_u / dtb / 1e9, 3), "unit": "GB/s",
                                 "api": "zb200_inflate_stream_host: the ONE raw stream above, runs found at its flush points, one batch"}
        L.zb200_host_free(C.c_void_p(h_back))
    # (3) against the reference on a prefix (rank 0, N = 1): bytes at levels >= 4, size + reference-decodes at levels 1-3
    if cpu is not None and g.rank == 0:
        npre = min(n, CPU_PREFIX)
        sec, outb, stream, ends = cpu.deflate(host, npre, level, strategy, cpu.threads, 1, want_stream=True)
        ends_gpu = d_end[:(npre + CHUNK - 1) // CHUNK].cpu().tolist()
        ours = ends_gpu[-1]
        # the reference's prefix stream ends in Z_FINISH (BFINAL) where ours continues with a sync marker
        last0 = ends_gpu[-2] if len(ends_gpu) > 1 else 0
        ours_body = bytes((C.c_uint8 * last0).from_address(h_out))
        e["size_vs_reference"] = round(ours / outb, 5) if npre < n else round(clen / outb, 5)
        gates["ratio_within_3pct"] = bool(e["size_vs_reference"] <= 1.03)
        if level >= 4:
            gates["bytes_identical_to_reference"] = bool(stream is not None and ends is not None and ours_body == stream.raw[:ends[-2] if len(ends) > 1 else 0]
                                                         and ends[:-1] == ends_gpu[:-1])
        if cpu.rp is not None:
            # the reference's inflate decodes OUR chunks (each a raw run that ends at a flush point) to the input bytes
            mem, prev = [], 0
            for i, en in enumerate(ends_gpu):
                mem.append((prev, en - prev, i * CHUNK, min(CHUNK, npre - i * CHUNK)))
                prev = en
            backbuf = C.create_string_buffer(npre)
            try:
                _, tot = cpu.inflate(h_out, mem, -15, cpu.threads, 1, C.addressof(backbuf), allow_open=True)
                gates["reference_inflate_decodes_our_stream"] = bool(tot == npre and backbuf.raw == C.string_at(host, npre))
            except RuntimeError as ex:
                gates["reference_inflate_decodes_our_stream"] = False
                gates["reference_inflate_error"] = str(ex)
        e["cpu_baseline"] = cpu_deflate_baseline(cpu, host, n, level, strategy)
    e["parity"] = gates
    L.zb200_host_free(C.c_void_p(h_out))
    kept = None
    if keep:
        kept = (host, d_in, n, lo)
    else:
        del d_in
        if own_host:
            L.zb200_host_free(C.c_void_p(host))
    del d_out
    torch.cuda.empty_cache()
    return e, kept


def inflate_leg(g, cpu, total, steps, warmup, plain=None, traffic=None):
    """C3 on this rank's shard.  `plain` = (pinned host ptr, device tensor, n, first byte) of markov text to reuse."""
    torch, zb, L = g.torch, g.zb, g.L
    lo, hi = shard(total, g.rank, g.world, MIB)
    n = hi - lo
    own = plain is None or plain[2] < n
    if own:
        host = host_alloc(L, n)
        fill(host, n, "markov", lo)
        d_plain = torch.empty(n, dtype=torch.uint8, device="cuda")
        d_plain.copy_(torch.frombuffer((C.c_uint8 * n).from_address(host), dtype=torch.uint8))
    else:
        host, d_plain = plain[0], plain[1][:n]
    sizes = member_sizes(n, SEED ^ (0x33 + g.rank))
    # members are made on the GPU, one deflate call per size class (level 6, gzip wrapper, Z_FINISH per member): at
    # level 6 that is the reference's member byte for byte, re-checked below on the CPU-baseline sample.
    # File order = schedule order; plain text of member i lies at the running sum of the sizes before it.
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += s
    by_size = {}
    for i, s in enumerate(sizes):
        by_size.setdefault(s, []).append(i)
    n_m = len(sizes)
    comp = [None] * n_m                                          # (class blob index, off, len)
    blobs = []
    d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    for s, idx in by_size.items():
        k = len(idx)
        # gather the class's plain text contiguously, compress it as k members of s bytes
        d_cls = torch.empty(k * s, dtype=torch.uint8, device="cuda")
        for j, i in enumerate(idx):
            d_cls[j * s:(j + 1) * s] = d_plain[offs[i]:offs[i] + s]
        cap = L.zb200_deflate_bound(k * s, s, zb.FRAME_GZIP_MEMBERS)
        d_blob = torch.empty(cap, dtype=torch.uint8, device="cuda")
        d_end = torch.zeros(k, dtype=torch.int64, device="cuda")
        r = L.zb200_deflate_dev(g.ctx.handle, d_cls.data_ptr(), k * s, s, 6, 0, zb.FRAME_GZIP_MEMBERS, 1, d_blob.data_ptr(), cap,
                                d_end.data_ptr(), d_tot.data_ptr(), g.sp)
        if r != 0:
            raise zb.ZB200Error(r, "zb200_deflate_dev(members)")
        torch.cuda.synchronize()
        ends = d_end.cpu().tolist()
        prev = 0
        for j, i in enumerate(idx):
            comp[i] = (len(blobs), prev, ends[j] - prev)
            prev = ends[j]
        blobs.append(d_blob[:prev].clone())
        del d_cls, d_blob
    # lay the members end to end in file order
    ctot = sum(c[2] for c in comp)
    d_file = torch.empty(ctot + 64, dtype=torch.uint8, device="cuda")
    members, ipos = [], 0
    for i in range(n_m):
        b, off, ln = comp[i]
        d_file[ipos:ipos + ln] = blobs[b][off:off + ln]
        members.append(zb.Member(ipos, ln, offs[i], sizes[i], 0, 0, 0))
        ipos += ln
    del blobs
    torch.cuda.empty_cache()
    arr = (zb.Member * n_m)(*members)
    d_members = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    d_res = torch.zeros(n_m * C.sizeof(zb.MemberResult), dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")

    def step():
        r = L.zb200_inflate_dev(g.ctx.handle, d_file.data_ptr(), d_out.data_ptr(), d_members.data_ptr(), n_m,
                                zb.WRAP_GZIP, 1, d_res.data_ptr(), g.sp)
        if r != 0:
            raise zb.ZB200Error(r, "zb200_inflate_dev")

    l0 = L.zb200_launch_count()
    ms = g.maxr(g.time_steps(step, steps, warmup))
    launches = (L.zb200_launch_count() - l0) * steps // (steps + warmup)
    kernels = g.kernel_shares(step)
    res = (zb.MemberResult * n_m).from_buffer_copy(d_res.cpu().numpy().tobytes())
    ok = all(r.status == 0 and r.out_len == sizes[i] for i, r in enumerate(res)) and bool(torch.equal(d_out, d_plain))
    tot_u, tot_c, tot_m = g.sumr(float(n), float(ctot), float(n_m))
    e = {"config": "C3", "value": round(tot_u / (ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms, 3), "steps": steps, "warmup": warmup,
         "total_bytes": int(tot_u), "bytes_per_gpu": n, "members": int(tot_m), "member_sizes": "log-uniform 64 KiB .. 1 MiB on a 65-value grid",
         "compressed_bytes": int(tot_c), "gpu_launches": int(launches), "kernels": kernels,
         "roofline": g.roofline(n + ctot, ms, kernels, traffic, "achieved/frac = (C + U) of this rank's shard / the whole step"),
         "parity": {"bit_exact_full_size": bool(ok)}}
    # ---- end to end: the multi-member file in pinned host memory -> pinned host output ----
    h_in, h_out = host_alloc(L, ctot + 64), host_alloc(L, n)
    if h_in and h_out:
        torch.frombuffer((C.c_uint8 * ctot).from_address(h_in), dtype=torch.uint8).copy_(d_file[:ctot])
        torch.cuda.synchronize()
        res2 = (zb.MemberResult * n_m)()

        def step_host():
            r = L.zb200_inflate_host(g.ctx.handle, C.c_void_p(h_in), C.c_void_p(h_out), arr, n_m, zb.WRAP_GZIP, 1, res2)
            if r != 0:
                raise zb.ZB200Error(r, "zb200_inflate_host")

        e2e_steps = max(2, min(steps, 3))
        dt = g.wall_steps(step_host, e2e_steps)
        ok2 = all(r.status == 0 for r in res2) and host_equals_device(torch, h_out, n, d_plain)
        e["e2e"] = {"value": round(tot_u / dt / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": int(tot_c), "d2h_bytes_per_step": int(tot_u),
                    "steps": e2e_steps, "api": "zb200_inflate_host, pinned host buffers, member table given, pieces pipelined"}
        e["parity"]["e2e_bit_exact_full_size"] = bool(ok2)
        # the same file with NO member table: member starts discovered on the device
        olen, nm, st = C.c_size_t(0), C.c_size_t(0), C.c_int(0)

        def step_scan():
            r = L.zb200_gunzip_host(g.ctx.handle, C.c_void_p(h_in), ctot, C.c_void_p(h_out), n, C.byref(olen), C.byref(st),
                                    None, 0, C.byref(nm))
            if r != 0:
                raise zb.ZB200Error(r, "zb200_gunzip_host")

        dt = g.wall_steps(step_scan, 2)
        ok3 = st.value == 0 and nm.value == n_m and olen.value == n and host_equals_device(torch, h_out, n, d_plain)
        e["e2e_no_index"] = {"value": round(tot_u / dt / 1e9, 3), "unit": "GB/s", "members_found": int(nm.value),
                             "api": "zb200_gunzip_host: no member table, member starts discovered on the device"}
        e["parity"]["no_index_bit_exact_full_size"] = bool(ok3)
        if cpu is not None and g.rank == 0:
            # the CPU baseline on the first members that make up >= 256 MiB, and the identity of those members with
            # the reference's own gzip members (deflateInit2(6, 15+16, 8, default) + Z_FINISH)
            k, acc = 0, 0
            while k < n_m and acc < CPU_PREFIX:
                acc += sizes[k]
                k += 1
            mem = [(members[i].in_off, members[i].in_len, offs[i], sizes[i]) for i in range(k)]
            back = C.create_string_buffer(acc)
            sec, tot = cpu.inflate(h_in, mem, 31, cpu.threads, 3, C.addressof(back))
            sec1, _ = cpu.inflate(h_in, mem[:max(1, k // 8)], 31, 1, 1, C.addressof(back))
            n1 = sum(m[3] for m in mem[:max(1, k // 8)])
            e["cpu_baseline"] = {"value": round(acc / sec / 1e9, 4), "unit": "GB/s", "cores": cpu.threads, "kind": cpu.kind,
                                 "sample": "first %d members (%d MiB of output), pthread pool (one z_stream per thread, inflateReset per member), best of 3" % (k, acc // MIB),
                                 "single_thread_value": round(n1 / sec1 / 1e9, 4)}
            e["parity"]["reference_inflate_agrees_on_sample"] = bool(tot == acc and back.raw == C.string_at(host, acc))
            theirs = cpu.deflate_members(host, [(offs[i], sizes[i]) for i in range(k)], 6, 31, cpu.threads)
            ident = sum(C.string_at(h_in + members[i].in_off, members[i].in_len) == theirs[i] for i in range(k))
            e["parity"]["members_identical_to_reference_deflate"] = bool(ident == k)
            e["members_checked_against_reference_deflate"] = k
    if h_in:
        L.zb200_host_free(C.c_void_p(h_in))
    if h_out:
        L.zb200_host_free(C.c_void_p(h_out))
    if own:
        L.zb200_host_free(C.c_void_p(host))
    del d_out, d_file, d_plain
    torch.cuda.empty_cache()
    return e


def checksum_leg(g, cpu, total, steps, warmup, traffic=None):
    """C2 on this rank's shard: fused CRC-32 + Adler-32, plus each alone; partials merged across ranks."""
    torch, zb, L = g.torch, g.zb, g.L
    lo, hi = shard(total, g.rank, g.world, 65536)
    n = hi - lo
    host = host_alloc(L, n)
    fill(host, n, "bytes", lo)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_in.copy_(torch.frombuffer((C.c_uint8 * n).from_address(host), dtype=torch.uint8))
    d_out2 = torch.zeros(2, dtype=torch.int32, device="cuda")

    def mk(which):
        def step():
            r = L.zb200_checksum_dev(g.ctx.handle, d_in.data_ptr(), n, which, 0, 1, d_out2.data_ptr(), g.sp)
            if r != 0:
                raise zb.ZB200Error(r, "zb200_checksum_dev")
        return step

    step = mk(zb.CRC32 | zb.ADLER32)
    l0 = L.zb200_launch_count()
    ms = g.maxr(g.time_steps(step, steps, warmup))
    launches = (L.zb200_launch_count() - l0) * steps // (steps + warmup)
    kernels = g.kernel_shares(step, 3)
    res = d_out2.cpu().numpy().astype("uint32")
    crc, adler = int(res[0]), int(res[1])
    tot = g.sumr(float(n))
    e = {"config": "C2", "value": round(tot / (ms * 1e-3) / 1e9, 2), "unit": "GB/s", "ms_per_step": round(ms, 4), "steps": steps, "warmup": warmup,
         "total_bytes": int(tot), "bytes_per_gpu": n, "which": "crc32+adler32 fused, one pass", "gpu_launches": int(launches), "kernels": kernels,
         "roofline": g.roofline(n, ms, kernels, traffic, "achieved/frac = N bytes read / the whole step")}
    for nm, w in (("crc32_only", zb.CRC32), ("adler32_only", zb.ADLER32)):
        ms1 = g.maxr(g.time_steps(mk(w), steps, 3))
        e[nm] = {"value": round(tot / (ms1 * 1e-3) / 1e9, 1), "unit": "GB/s", "ms_per_step": round(ms1, 4),
                 "roofline_frac": round(n / (ms1 * 1e-3) / 1e9 / g.peak, 4)}
    c_crc, c_adler = C.c_uint32(0), C.c_uint32(0)

    def step_host():
        r = L.zb200_checksum_host(g.ctx.handle, C.c_void_p(host), n, zb.CRC32 | zb.ADLER32, 0, 1, C.byref(c_crc), C.byref(c_adler))
        if r != 0:
            raise zb.ZB200Error(r, "zb200_checksum_host")

    e2e_steps = max(2, min(steps, 5))
    dt = g.wall_steps(step_host, e2e_steps)
    e["e2e"] = {"value": round(tot / dt / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": int(tot), "d2h_bytes_per_step": 8 * g.world,
                "steps": e2e_steps, "api": "zb200_checksum_host, pinned host buffer"}
    gates = {"host_path_equals_device_path": bool((c_crc.value, c_adler.value) == (crc, adler))}
    # ---- across ranks: the combined value of all shards == the reference's value of the logical buffer ----
    # every rank runs the reference over its own shard on its share of the host cores; the partials of both sides are
    # folded in rank order — ours with the library's combine, the reference's with the reference's own crc32_combine.
    from zlib_wasm_b200 import shard as zshard
    job_crc, job_adler, job_n = zshard.gather_checksums(g.dist, crc, adler, n, device="cuda")
    e["checks"] = {"crc32": "%08x" % job_crc, "adler32": "%08x" % job_adler, "bytes": int(job_n)}
    if cpu is not None:
        thr = max(1, cpu.threads // g.world)
        sec, rc, ra = cpu.checksum(host, n, thr, 3 if g.world == 1 else 1)
        if g.world > 1:
            mine = torch.tensor([rc, ra, n], dtype=torch.int64, device="cuda")
            got = [torch.zeros_like(mine) for _ in range(g.world)]
            g.dist.all_gather(got, mine)
            parts = [tuple(int(x) for x in t.tolist()) for t in got]
        else:
            parts = [(rc, ra, n)]
        ref_crc, ref_adler = cpu.combine(parts)
        gates["combined_over_ranks_equals_reference"] = bool((ref_crc, ref_adler) == (job_crc, job_adler))
        if g.world == 1 and g.rank == 0:
            n1 = min(n, 256 * MIB)
            sec1, _, _ = cpu.checksum(host, n1, 1, 1)
            e["cpu_baseline"] = {"value": round(n / sec / 1e9, 3), "unit": "GB/s", "cores": cpu.threads, "kind": cpu.kind,
                                 "sample": "the whole %d MiB buffer, crc32_z + adler32_z, one range per thread + crc32_combine / adler32_combine, best of 3" % (n // MIB),
                                 "single_thread_value": round(n1 / sec1 / 1e9, 3)}
    e["parity"] = gates
    L.zb200_host_free(C.c_void_p(host))
    del d_in
    torch.cuda.empty_cache()
    return e


def small_call_leg(cpu):
    """Latency of small calls through the zlib.h surface itself (pageable host buffers, one host thread), this library next
    to the reference: crc32, compress2 (level 6), uncompress, and deflate() fed in 16 KiB slices (zlib_deflate_process /
    zpipe style).  Median of 15 calls after 3 warm-up calls, microseconds."""
    import refz
    import zlib_wasm_b200 as zb
    libs = {"b200": refz.ZlibBinding(zb.LIB_PATH, "")}
    if cpu is not None and cpu.ref is not None:
        libs["reference"] = cpu.ref
    out = {"unit": "us per call (median of 15)", "sizes": {}}

    def med(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(15):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        return round(ts[len(ts) // 2] * 1e6, 1)

    for n in (64, 1024, 16384, 262144, 1048576):
        d = refz.gen(n, refz.GEN_TEXT, seed=n)
        row = {}
        for name, z in libs.items():
            cap = z.compressBound(n) + 64
            dst, back = C.create_string_buffer(cap), C.create_string_buffer(n)
            dl = C.c_ulong(cap)
            z.compress2(dst, C.byref(dl), d, n, 6)
            comp = dst.raw[:dl.value]

            def f_comp():
                l = C.c_ulong(cap)
                z.compress2(dst, C.byref(l), d, n, 6)

            def f_unc():
                l = C.c_ulong(n)
                z.uncompress(back, C.byref(l), comp, len(comp))

            def f_stream():
                z.deflate_stream(d, 6, 0, refz.WRAP_ZLIB, 0, in_slice=16384, out_slice=16384)

            row[name] = {"crc32": med(lambda: z.crc32(0, d, n)), "compress2_l6": med(f_comp), "uncompress": med(f_unc),
                         "deflate_16k_slices": med(f_stream)}
        out["sizes"][str(n)] = row
    return out


def link_probe(g, nbytes=1 << 30, reps=4):
    """What the host link gives THIS box with every rank copying at once and no kernel running: pinned host memory (placed
    on the GPU's NUMA node) -> device, device -> host, and both directions together.  The end-to-end legs cannot beat
    these; N ranks sharing PCIe switches / memory controllers show here, not in the kernels."""
    torch, L = g.torch, g.L
    dev = torch.cuda.current_device()
    host = host_alloc(L, 2 * nbytes, dev)
    h = torch.frombuffer((C.c_uint8 * (2 * nbytes)).from_address(host), dtype=torch.uint8)
    h[::4096] = 1                                            # (every page touched)
    d = torch.empty(2 * nbytes, dtype=torch.uint8, device="cuda")
    s2 = torch.cuda.Stream()
    out = {"bytes_per_copy": nbytes, "copies": reps, "numa_node": gpu_numa(dev)[0],
           "local_cpus": len(gpu_numa(dev)[1]) if gpu_numa(dev)[1] else None, "host_cpus": len(os.sched_getaffinity(0))}

    def run(kind):
        def once():
            if kind in ("h2d", "both"):
                d[:nbytes].copy_(h[:nbytes], non_blocking=True)
            if kind == "d2h":
                h[nbytes:].copy_(d[nbytes:], non_blocking=True)
            if kind == "both":
                with torch.cuda.stream(s2):
                    h[nbytes:].copy_(d[nbytes:], non_blocking=True)
        once()
        torch.cuda.synchronize()
        g.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mine = nbytes * reps * (2 if kind == "both" else 1) / dt / 1e9
        worst = g.maxr(dt)
        return round(mine, 2), round(g.world * nbytes * reps * (2 if kind == "both" else 1) / worst / 1e9, 2)

    for kind in ("h2d", "d2h", "both"):
        mine, agg = run(kind)
        out[kind + "_GBps_this_rank"] = mine
        out[kind + "_GBps_all_ranks"] = agg
    del d, h
    L.zb200_host_free(C.c_void_p(host))
    torch.cuda.empty_cache()
    return out


def single_member_leg(g, cpu, total=256 * MIB):
    """ONE zlib stream without flush points (what compress2 / gzip write), the reference's bytes: decoded through
    zb200_inflate_stream_host (chunks between dynamic block headers in parallel, csrc/zb_inflate_blocks.cuh) from pinned
    host memory, next to the reference's inflate of the same stream on one core (a stream is serial for it)."""
    import zlib
    torch, L = g.torch, g.L
    n = total
    host = host_alloc(L, n)
    fill(host, n, "markov", 0)
    raw = C.string_at(host, n)
    t0 = time.perf_counter()
    comp = zlib.compress(raw, 6)                              # (system zlib 1.3: the reference's bytes)
    t_comp = time.perf_counter() - t0
    h_in, h_out = host_alloc(L, len(comp) + 64), host_alloc(L, n + 64)
    C.memmove(h_in, comp, len(comp))
    res = g.zb.MemberResult()

    def call():
        r = L.zb200_inflate_stream_host(g.ctx.handle, C.c_void_p(h_in), len(comp), g.zb.WRAP_ZLIB, C.c_void_p(h_out), n + 64, C.byref(res))
        if r != 0 or res.status != 0 or res.out_len != n:
            raise RuntimeError("single member: r=%d status=%d out_len=%d" % (r, res.status, res.out_len))

    call()
    ok = C.string_at(h_out, n) == raw
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        call()
        ts.append(time.perf_counter() - t0)
    kern = g.kernel_shares(call, 1)
    e = {"config": "one member of %d MiB (zlib level 6, no flush points), pinned host buffers" % (n >> 20), "compressed_bytes": len(comp),
         "e2e": {"value": round(n / min(ts) / 1e9, 3), "unit": "GB/s", "ms": round(min(ts) * 1e3, 2), "h2d_bytes_per_step": len(comp),
                 "d2h_bytes_per_step": n, "api": "zb200_inflate_stream_host"},
         "device_ms": round(sum(v["ms_per_step"] for v in kern.values()), 3), "kernels": kern, "parity": {"bit_exact": bool(ok)},
         "reference_compress_one_core_s": round(t_comp, 2)}
    if cpu is not None and cpu.ref is not None:
        import refz
        back = C.create_string_buffer(n)
        bl = C.c_ulong(n)
        t0 = time.perf_counter()
        rc = cpu.ref.uncompress(back, C.byref(bl), comp, len(comp))
        dt = time.perf_counter() - t0
        e["cpu_baseline"] = {"value": round(n / dt / 1e9, 3), "unit": "GB/s", "cores": 1, "kind": "reference",
                             "sample": "the same stream through the reference's uncompress(): one stream is one core's work", "rc": int(rc)}
    for p in (host, h_in, h_out):
        L.zb200_host_free(C.c_void_p(p))
    return e


def chunk_carry_leg(g, cpu, total=512 * MIB):
    """SURVEY 8 f3, carry-over: the same chunks compressed behind the 32 KiB before them (zb200.h ZB200_CHUNK_CARRY — per chunk
    what deflateSetDictionary(previous 32 KiB) + deflate(Z_SYNC_FLUSH) gives, the pigz scheme) next to the independent
    Z_FULL_FLUSH chunks of the headline: size, end-to-end and kernel time of both through zb200_deflate_host on pinned buffers;
    gates: the reference's uncompress() and this library's one-stream decoder give the input back."""
    L, zb = g.L, g.zb
    n = total
    cap = L.zb200_deflate_bound(n, CHUNK, zb.FRAME_ZLIB)
    h_out, h_back = host_alloc(L, cap), host_alloc(L, n + 64)
    out = {"bytes": n, "chunk": CHUNK, "levels": {}}
    ok_ref = ok_own = True
    for level, gen in ((1, "markov"), (6, "mixed")):
        host = host_alloc(L, n)
        fill(host, n, gen, 0)
        row = {"generator": gen}
        for key, flag in (("independent", 0), ("carried", zb.CHUNK_CARRY)):
            olen = C.c_size_t(cap)

            def call():
                olen.value = cap
                r = L.zb200_deflate_host(g.ctx.handle, C.c_void_p(host), n, CHUNK, level, 0, zb.FRAME_ZLIB | flag, 1, C.c_void_p(h_out), C.byref(olen), None, None)
                if r != 0:
                    raise RuntimeError("chunk_carry level %d: %s" % (level, zb.last_error()))

            call()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                call()
                ts.append(time.perf_counter() - t0)
            kern = g.kernel_shares(call, 1)
            row[key] = {"compressed_bytes": olen.value, "e2e_ms": round(min(ts) * 1e3, 2), "e2e_GBps": round(n / min(ts) / 1e9, 2),
                        "kernel_ms": round(sum(v["ms_per_step"] for v in kern.values()), 2)}
        # (h_out now holds the carried stream)
        res = zb.MemberResult()
        t0 = time.perf_counter()
        r = L.zb200_inflate_stream_host(g.ctx.handle, C.c_void_p(h_out), row["carried"]["compressed_bytes"], zb.WRAP_ZLIB, C.c_void_p(h_back), n + 64, C.byref(res))
        row["own_decode_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        own = r == 0 and res.status == 0 and res.out_len == n and C.string_at(h_back, n) == C.string_at(host, n)
        ok_own = ok_own and own
        if cpu is not None and cpu.ref is not None:
            C.memset(h_back, 0, n)
            bl = C.c_ulong(n)
            rc = cpu.ref.uncompress(C.c_void_p(h_back), C.byref(bl), C.c_void_p(h_out), row["carried"]["compressed_bytes"])
            ok_ref = ok_ref and rc == 0 and bl.value == n and C.string_at(h_back, n) == C.string_at(host, n)
        row["size_carried_vs_independent"] = round(row["carried"]["compressed_bytes"] / row["independent"]["compressed_bytes"], 5)
        row["kernel_time_carried_vs_independent"] = round(row["carried"]["kernel_ms"] / row["independent"]["kernel_ms"], 4)
        out["levels"][str(level)] = row
        L.zb200_host_free(C.c_void_p(host))
    out["parity"] = {"own_stream_decoder_round_trip": bool(ok_own)}
    if cpu is not None and cpu.ref is not None:
        out["parity"]["reference_inflates_carried_stream"] = bool(ok_ref)
    for p in (h_out, h_back):
        L.zb200_host_free(C.c_void_p(p))
    return out


def one_shot_leg(cpu, n=64 * MIB):
    """compress2() / uncompress() of one buffer through the zlib.h surface itself (pageable host buffers, one host thread; the
    reference's API is serial), this library next to the reference: level 6 and 9, wall time of the calls, and whether the
    two libraries' streams are the same bytes (levels 4-9 emit the reference's one run of blocks whatever the length)."""
    import refz
    import zlib_wasm_b200 as zb
    libs = {"b200": refz.ZlibBinding(zb.LIB_PATH, "")}
    if cpu is not None and cpu.ref is not None:
        libs["reference"] = cpu.ref
    d = refz.gen(n, refz.GEN_TEXT, seed=0x9E37)
    out = {"bytes": n, "generator": "word text", "unit": "ms per call", "levels": {}}
    for level in (6, 9):
        row, streams = {}, {}
        for name, z in libs.items():
            cap = z.compressBound(n) + 64
            dst, back = C.create_string_buffer(cap), C.create_string_buffer(n)
            best_c = best_u = None
            for rep in range(2 if name == "b200" else 1):         # (the library's device buffers grow on the first call)
                dl = C.c_ulong(cap)
                t0 = time.perf_counter()
                rc = z.compress2(dst, C.byref(dl), d, n, level)
                t1 = time.perf_counter()
                bl = C.c_ulong(n)
                ru = z.uncompress(back, C.byref(bl), dst, dl.value)
                t2 = time.perf_counter()
                if rc != 0 or ru != 0 or bl.value != n:
                    raise RuntimeError("one_shot %s level %d: compress2 %d uncompress %d" % (name, level, rc, ru))
                best_c = t1 - t0 if best_c is None else min(best_c, t1 - t0)
                best_u = t2 - t1 if best_u is None else min(best_u, t2 - t1)
            streams[name] = dst.raw[:dl.value]
            row[name] = {"compress2": round(best_c * 1e3, 2), "uncompress": round(best_u * 1e3, 2), "compressed_bytes": dl.value,
                         "round_trip_ok": bool(back.raw == d)}
        if len(streams) == 2:
            row["same_bytes_as_reference"] = bool(streams["b200"] == streams["reference"])
        out["levels"][str(level)] = row
    out["parity"] = {"round_trip": all(r[k]["round_trip_ok"] for r in out["levels"].values() for k in r if isinstance(r[k], dict)),
                     "streams_identical_to_reference": all(r.get("same_bytes_as_reference", True) for r in out["levels"].values())}
    return out
