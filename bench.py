#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its configuration C2 (CRC-32 + Adler-32
over a 4 GiB synthetic buffer per GPU, per-GPU folding + crc32_combine across
GPUs), plus bounded side measurements of the other hot-path legs (inflate of
multi-member gzip, deflate L1 / L6 over 256 KiB Z_FULL_FLUSH chunks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  See the task contract for the keys.  A "step"
is one fused CRC-32+Adler-32 pass over the rank's resident 4 GiB shard.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GIB = 1 << 30
SEED = 0x9E3779B97F4A7C15
METRIC = "crc32_adler32_GBps"
UNIT = "GB/s"


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full
    capture (profiles/*ck_big_ncu_full.txt), or None."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ck_big_ncu_full.txt"))):
        rd = wr = None
        for line in open(path):
            m = re.search(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m:
                v = float(m.group(2)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), 1)
                if m.group(1) == "read":
                    rd = v
                else:
                    wr = v
            if rd is not None and wr is not None:
                best = int(rd + wr)
                break
    return best


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference's crc32_z + adler32_z
# (oracle/_ref/libzref.so), chunk-parallel over host threads, partials merged
# with the reference's own crc32_combine / adler32_combine.
# ---------------------------------------------------------------------------
def cpu_checksums(ref, buf_addr, n, threads):
    import concurrent.futures as cf
    piece = (n + threads - 1) // threads
    piece = (piece + 63) & ~63
    spans = [(o, min(piece, n - o)) for o in range(0, n, piece)]

    def work(span):
        o, k = span
        return ref.crc32_z(0, buf_addr + o, k), ref.adler32_z(1, buf_addr + o, k), k

    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, spans))
    crc, adler = parts[0][0], parts[0][1]
    for c, a, k in parts[1:]:
        crc = ref.crc32_combine(crc, c, k)
        adler = ref.adler32_combine(adler, a, k)
    return crc, adler


def load_cpu_ref():
    import refz
    if refz.have_ref():
        return refz.ref(), "reference"
    return None, "port"


def run_reference_arm(args):
    """--impl reference: rank 0 only; CPU reference on all host threads."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import refz
    ref, kind = load_cpu_ref()
    threads = host_threads()
    sample = min(GIB, args.bytes)
    data = refz.gen(sample, refz.GEN_BYTES, SEED)
    buf = C.create_string_buffer(data, sample)
    addr = C.addressof(buf)
    if ref is None:
        o = refz.oracle()

        class Port:
            crc32_z = staticmethod(lambda c, p, k: o.c_crc32(c, p, k))
            adler32_z = staticmethod(lambda a, p, k: o.c_adler32(a, p, k))
            crc32_combine = staticmethod(lambda a, b, k: o.c_crc32_combine(a, b, k))
            adler32_combine = staticmethod(lambda a, b, k: o.c_adler32_combine(a, b, k))
        ref = Port
    for _ in range(args.warmup):
        cpu_checksums(ref, addr, sample, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_checksums(ref, addr, sample, threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "bytes4g_per_gpu", "note": "CPU reference (zlib 1.3.1.1-motley crc32_z+adler32_z), "
                       "chunk-parallel over host threads + crc32_combine/adler32_combine; each step a %d-byte sample" % sample},
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "%d bytes of bytes4g (50%% word text / 50%% xorshift), per step" % sample},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bytes", type=int, default=4 * GIB, help="bytes per GPU (config C2: 4 GiB)")
    ap.add_argument("--no-extras", action="store_true", help="skip the inflate / deflate side measurements")
    ap.add_argument("--extras-mib", type=int, default=512, help="uncompressed MiB used by each side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    # torchrun pins OMP_NUM_THREADS=1; the synthetic-data generator (OpenMP, not part of the
    # timed region) may use this rank's share of the host cores
    os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // max(1, env_int("WORLD_SIZE", 1))))
    # Libraries (NCCL's version banner, torch warnings) may write to stdout; the contract is
    # ONE JSON line there.  Route fd 1 to stderr for the run and keep the real stdout aside.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import zlib_wasm_b200 as zb
    import refz

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = zb.lib()
    ctx = zb.Context(local)
    # a real (non-default) stream: the C ABI treats a NULL stream as "the context's
    # own stream", and CUDA events only see the stream they are recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert sp.value, "expected a non-default stream handle"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: this rank's shard of the logical (4 GiB x world) buffer ----
    n = args.bytes
    blocks_per_rank = (n + 65535) // 65536
    h_in = L.zb200_host_alloc(n)                      # pinned: e2e DMA-s straight out of it
    if not h_in:
        raise SystemExit("pinned host allocation of %d bytes failed" % n)
    zg = C.CDLL(os.path.join(ROOT, "tools", "libzgen.so"))
    zg.zgen_fill.restype = None
    zg.zgen_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64]
    zg.zgen_fill(h_in, n, refz.GEN_BYTES, SEED, rank * blocks_per_rank)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    h_view = torch.frombuffer((C.c_uint8 * n).from_address(h_in), dtype=torch.uint8)
    d_in.copy_(h_view)
    torch.cuda.synchronize()
    d_out2 = torch.zeros(2, dtype=torch.int32, device="cuda")
    which = zb.CRC32 | zb.ADLER32

    def step_dev():
        r = L.zb200_checksum_dev(ctx.handle, d_in.data_ptr(), n, which, 0, 1, d_out2.data_ptr(), sp)
        if r != 0:
            raise zb.ZB200Error(r, "zb200_checksum_dev")

    # ---- device-resident timing (value, roofline) ----
    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = L.zb200_launch_count()
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0.record(stream)
    for a, b in ev:
        a.record(stream)
        step_dev()
        b.record(stream)
    t_all1.record(stream)
    barrier()
    launches = L.zb200_launch_count() - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    per_launch_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    res = d_out2.cpu().numpy().astype("uint32")
    crc, adler = int(res[0]), int(res[1])

    # ---- end to end through the host-buffer C ABI (H2D inside) ----
    c_crc, c_adler = C.c_uint32(0), C.c_uint32(0)

    def step_host():
        r = L.zb200_checksum_host(ctx.handle, h_in, n, which, 0, 1, C.byref(c_crc), C.byref(c_adler))
        if r != 0:
            raise zb.ZB200Error(r, "zb200_checksum_host")

    e2e_steps = max(2, min(args.steps, 5))
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    if (c_crc.value, c_adler.value) != (crc, adler):
        raise SystemExit("host-path and device-path checksums disagree: %x/%x vs %x/%x" % (c_crc.value, c_adler.value, crc, adler))

    # ---- max over ranks, combine across ranks ----
    if world > 1:
        t = torch.tensor([total_ms, per_launch_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, per_launch_ms, e2e_s = [float(x) for x in t.cpu()]
        parts = torch.zeros(world, 2, dtype=torch.int64, device="cuda")
        mine = torch.tensor([crc, adler], dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(parts.view(-1), mine)
        parts = parts.cpu().tolist()
    else:
        parts = [[crc, adler]]
    job_crc, job_adler = parts[0]
    for c, a in parts[1:]:                            # host-side crc32_combine / adler32_combine (crc32.c:1021, adler32.c:133)
        job_crc = L.zb200_crc32_combine(job_crc, c, n)
        job_adler = L.zb200_adler32_combine(job_adler, a, n)

    line = None
    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * n / (ms_per_step * 1e-3) / 1e9
        peak, peak_src = peaks()
        achieved = n / (per_launch_ms * 1e-3) / 1e9
        # CPU baseline: the compiled reference on the host cores, bounded sample
        ref, kind = load_cpu_ref()
        threads = host_threads()
        sample = min(n, GIB)
        cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": kind, "sample": "skipped"}
        if ref is not None:
            t0 = time.perf_counter()
            rc, ra = cpu_checksums(ref, h_in, sample, threads)
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            one = min(sample, 256 << 20)
            ref.crc32_z(0, h_in, one), ref.adler32_z(1, h_in, one)
            dt1 = time.perf_counter() - t1
            cpu = {"value": round(sample / dt / 1e9, 3), "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": "first %d bytes of rank 0's shard, crc32_z+adler32_z chunk-parallel + combine" % sample,
                   "single_thread_value": round(one / dt1 / 1e9, 3)}
            # parity gate of the run itself: GPU prefix checksum == reference prefix checksum
            r = L.zb200_checksum_dev_sync(ctx.handle, d_in.data_ptr(), sample, which, 0, 1, C.byref(c_crc), C.byref(c_adler), sp)
            if r != 0 or (c_crc.value, c_adler.value) != (rc, ra):
                raise SystemExit("parity failure against the reference on the bench input")
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "bytes4g_per_gpu", "bytes_per_gpu": n, "which": "crc32+adler32 fused, one pass",
                       "l2": "input (4 GiB) is far larger than the 126 MB L2; no flush needed",
                       "checks": {"crc32": "%08x" % job_crc, "adler32": "%08x" % job_adler}},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": profiled_traffic(), "peak_source": peak_src,
                         "kernel": "ck_big_kernel<crc,adler>", "algorithmic_bytes_per_launch": n},
            "cpu_baseline": cpu,
            "e2e": {"value": round(world * n / e2e_s / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": 8,
                    "api": "zb200_checksum_host on pinned host memory", "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
    # ---- side measurements (rank 0 prints; every rank runs its shard) ----
    if not args.no_extras:
        try:
            import bench_extras
            extra = {}
            # the two checksums on their own ("report both separate and fused"), same 4 GiB shard
            for name, w in (("crc32_only", zb.CRC32), ("adler32_only", zb.ADLER32)):
                def one(w=w):
                    r = L.zb200_checksum_dev(ctx.handle, d_in.data_ptr(), n, w, 0, 1, d_out2.data_ptr(), sp)
                    if r != 0:
                        raise zb.ZB200Error(r, "zb200_checksum_dev")
                for _ in range(3):
                    one()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                a.record(stream)
                for _ in range(args.steps):
                    one()
                b.record(stream)
                barrier()
                ms1 = a.elapsed_time(b) / args.steps
                if world > 1:
                    t = torch.tensor([ms1], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms1 = float(t.item())
                pk, _src = peaks()
                extra[name] = {"value": round(world * n / (ms1 * 1e-3) / 1e9, 1), "unit": UNIT, "ms_per_step": round(ms1, 4),
                               "roofline_frac": round(n / (ms1 * 1e-3) / 1e9 / pk, 4)}
            del d_in                                   # free the 4 GiB shard before the deflate / inflate workloads
            torch.cuda.empty_cache()
            extra.update(bench_extras.run(ctx, rank, world, args.extras_mib << 20, barrier))
            if line is not None:
                line["extra"] = extra
        except Exception as e:  # a side measurement must never take the headline down
            if line is not None:
                line["extra"] = {"error": repr(e)}
    if line is not None:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    L.zb200_host_free(h_in)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
