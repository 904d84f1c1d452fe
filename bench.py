#!/usr/bin/env python
"""bench.py — BASELINE.json's metric, "deflate L1/L6 + inflate GB/s at 1-8 B200; CRC32 GB/s vs HBM roofline",
every leg at its configuration's size, in ONE JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Headline (top-level keys) = config C4, north_star's first target: deflate level 1 over markov8g (8 GiB of
order-1 word text, 256 KiB Z_FULL_FLUSH chunks), STRONG scaling — the 8 GiB are split evenly over the N ranks,
chunk c -> rank floor(c*N/chunks), no data-path collective.  A "step" is one pass of the whole deflate pipeline
over the rank's resident shard.  `legs` carries the other configurations with the same keys each:
C5 (deflate L6 / L9 x default / Z_FILTERED on mixed2g), C3 (inflate of 8 GiB of gzip members) and C2 (CRC-32 +
Adler-32 over 4 GiB); see bench_legs.py.

--impl reference times the UNMODIFIED reference (oracle/_ref, compiled in place from /root/reference) on the
host's cores through the same pthread pool, leg for leg, each step a bounded sample of the same bytes.
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

import bench_legs as BL  # noqa: E402

METRIC = "deflate_l1_GBps"
UNIT = "GB/s"


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


def headline_config(total, world):
    """The `config` object — identical in both arms (the driver compares them)."""
    return {"workload": "C4 markov8g: deflate level 1 (deflate_fast), Z_DEFAULT_STRATEGY, raw stream of 256 KiB Z_FULL_FLUSH chunks",
            "total_bytes": int(total), "chunk": BL.CHUNK, "level": 1, "strategy": 0, "generator": "order-1 word text, seed 0x9E3779B97F4A7C15",
            "l2": "inputs far larger than the 126 MB L2 (>= 1 GiB per rank and step); no flush needed"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region of the headline."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.th = None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=3)

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


# ---------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """--impl reference: rank 0 only; the reference's CPU implementation on all host threads, every leg of the
    metric on a bounded sample (>= 256 MiB prefix) of the SAME bytes the GPU arm compresses."""
    if env_int("RANK", 0) != 0:
        return
    cpu = BL.Cpu()
    thr = cpu.threads
    total = BL.scaled(BL.DEFLATE_LEGS["deflate_l1"][1], args.scale)
    sample = min(total, BL.CPU_PREFIX)
    buf = C.create_string_buffer(sample)
    addr = C.addressof(buf)
    BL.fill(addr, sample, "markov", 0)
    for _ in range(args.warmup):
        cpu.deflate(addr, sample, 1, 0, thr, 1)
    t0 = time.perf_counter()
    outb = 0
    for _ in range(args.steps):
        _, outb, _, _ = cpu.deflate(addr, sample, 1, 0, thr, 1)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt / 1e9
    legs = {}
    if not args.no_legs:
        # C3: the reference inflates the reference's own members of the same text
        sizes = BL.member_sizes(total, BL.SEED ^ 0x33)
        k, acc = 0, 0
        while k < len(sizes) and acc < sample:
            acc += sizes[k]
            k += 1
        acc = min(acc, sample)
        spans, o = [], 0
        for i in range(k):
            s = min(sizes[i], sample - o)
            if s <= 0:
                break
            spans.append((o, s))
            o += s
        blob = cpu.deflate_members(addr, spans, 6, 31, thr)
        mem, off = [], 0
        for (so, sl), z in zip(spans, blob):
            mem.append((off, len(z), so, sl))
            off += len(z)
        blob = b"".join(blob)
        bb = C.create_string_buffer(blob, len(blob))
        back = C.create_string_buffer(o)
        sec, tot = cpu.inflate(C.addressof(bb), mem, 31, thr, 3, C.addressof(back))
        legs["inflate"] = {"config": "C3", "value": round(o / sec / 1e9, 4), "unit": UNIT, "cores": thr, "kind": cpu.kind,
                           "sample": "first %d members (%d MiB of output) of the GPU arm's rank-0 schedule, best of 3" % (len(mem), o >> 20),
                           "bit_exact": bool(tot == o and back.raw == C.string_at(addr, o))}
        del bb, back, blob
        # C5
        mixed = C.create_string_buffer(sample)
        BL.fill(C.addressof(mixed), sample, "mixed", 0)
        for name, (kind, tot_b, level, strategy) in BL.DEFLATE_LEGS.items():
            if name == "deflate_l1":
                legs[name] = {"config": "C4", "value": round(val, 4), "unit": UNIT, "cores": thr, "kind": cpu.kind, "ratio": round(sample / outb, 4),
                              "sample": "the headline"}
                continue
            sec, ob, _, _ = cpu.deflate(C.addressof(mixed), sample, level, strategy, thr, 3)
            legs[name] = {"config": "C5", "value": round(sample / sec / 1e9, 4), "unit": UNIT, "cores": thr, "kind": cpu.kind,
                          "ratio": round(sample / ob, 4), "sample": "first %d MiB of mixed2g, best of 3" % (sample >> 20)}
        del mixed
        # C2: the whole 4 GiB buffer
        n2 = BL.scaled(BL.CHECKSUM_TOTAL, args.scale)
        b2 = C.create_string_buffer(n2)
        BL.fill(C.addressof(b2), n2, "bytes", 0)
        sec, crc, adler = cpu.checksum(C.addressof(b2), n2, thr, 3)
        legs["checksum"] = {"config": "C2", "value": round(n2 / sec / 1e9, 3), "unit": UNIT, "cores": thr, "kind": cpu.kind,
                            "sample": "the whole %d MiB buffer, best of 3" % (n2 >> 20), "checks": {"crc32": "%08x" % crc, "adler32": "%08x" % adler}}
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": headline_config(total, args.gpus),
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": thr, "kind": cpu.kind,
                             "sample": "each step = the first %d MiB of markov8g, 256 KiB Z_FULL_FLUSH chunks, pthread pool of %d threads "
                                       "(one z_stream per thread, deflateReset per chunk)" % (sample >> 20, thr), "ratio": round(sample / outb, 4)},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "legs": legs}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def dram_traffic_probe(timeout=240):
    """DRAM bytes per input byte of each leg's kernels, measured now: a bounded pass of every leg
    (tools/traffic_probe.py) under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`.  Byte counts,
    not timings — nothing timed here is reported.  Returns {leg: {...}} or {"error": ...}."""
    import shutil
    import tempfile
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {"error": "ncu not found"}
    tmp = tempfile.mkdtemp(prefix="zb_traffic_")
    csv, side = os.path.join(tmp, "t.csv"), os.path.join(tmp, "side.json")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv", "--log-file", csv,
           sys.executable, os.path.join(ROOT, "tools", "traffic_probe.py"), side]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0")))
        if p.returncode != 0 or not os.path.exists(side) or not os.path.exists(csv):
            return {"error": "ncu probe failed rc=%d: %s" % (p.returncode, (p.stderr or p.stdout)[-300:])}
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import traffic_probe
        return traffic_probe.attribute(csv, side)
    except Exception as ex:  # a probe must never take the bench down
        return {"error": repr(ex)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("ZB200_BENCH_SCALE", "1")),
                    help="shrink every configuration by this factor (testing only; 1 = BASELINE.json's sizes)")
    ap.add_argument("--no-legs", action="store_true", help="headline only")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu DRAM-traffic probe")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    # torchrun pins OMP_NUM_THREADS=1; the synthetic-data generator (OpenMP, outside every timed region) may use
    # this rank's share of the host cores
    world = env_int("WORLD_SIZE", 1)
    os.environ["OMP_NUM_THREADS"] = str(max(1, BL.host_threads() // max(1, world)))
    # Libraries (NCCL's banner, torch warnings) may write to stdout; the contract is ONE JSON line there.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import zlib_wasm_b200 as zb

    rank, local = env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = zb.Context(local)
    # a real (non-default) stream: the C ABI treats NULL as "the context's own stream", and CUDA events only see the
    # stream they are recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g = BL.Gpu(torch, dist, zb, ctx, stream, rank, world)
    cpu = BL.Cpu() if world == 1 else None            # cpu_baseline: rank 0 at N = 1 only
    cpu_gate = BL.Cpu()                               # ... but the cross-rank checksum gate runs the reference on every rank
    t_start = time.time()

    traffic = {}
    if world == 1 and not args.no_traffic and not os.environ.get("ZB200_BENCH_NO_NCU"):
        traffic = dram_traffic_probe()

    def tr(leg, nbytes):
        t = traffic.get(leg) if isinstance(traffic, dict) else None
        return int(t["dram_bytes_per_input_byte"] * nbytes) if t and "dram_bytes_per_input_byte" in t else None

    # ---- headline: C4 ----
    total = BL.scaled(BL.DEFLATE_LEGS["deflate_l1"][1], args.scale)
    lo, hi = BL.shard(total, rank, world, BL.CHUNK)
    clocks = ClockSampler(local)
    head, kept = BL.deflate_leg(g, cpu, "deflate_l1", total, args.steps, args.warmup, keep=True, traffic=tr("deflate_l1", hi - lo), clocks=clocks)
    legs = {"deflate_l1": head}
    errors = {}
    if not args.no_legs:
        leg_steps = max(2, min(args.steps, 5))
        try:                                            # C3 reuses the resident markov text
            legs["inflate"] = BL.inflate_leg(g, cpu, BL.scaled(BL.INFLATE_TOTAL, args.scale), leg_steps, 3, plain=kept,
                                             traffic=tr("inflate", hi - lo))
        except Exception as ex:
            errors["inflate"] = repr(ex)
    if kept:
        zb.lib().zb200_host_free(C.c_void_p(kept[0]))
        kept = None
        torch.cuda.empty_cache()
    if not args.no_legs:
        for name in ("deflate_l6", "deflate_l6_filtered", "deflate_l9", "deflate_l9_filtered"):
            try:
                tot5 = BL.scaled(BL.DEFLATE_LEGS[name][1], args.scale)
                l5, h5 = BL.shard(tot5, rank, world, BL.CHUNK)
                st = leg_steps if "l6" in name else 2
                legs[name], _ = BL.deflate_leg(g, cpu, name, tot5, st, 3, traffic=tr(name, h5 - l5))
            except Exception as ex:
                errors[name] = repr(ex)
        try:
            tot2 = BL.scaled(BL.CHECKSUM_TOTAL, args.scale)
            l2, h2 = BL.shard(tot2, rank, world, 65536)
            legs["checksum"] = BL.checksum_leg(g, cpu_gate, tot2, max(leg_steps, min(args.steps, 20)), 3, traffic=tr("checksum", h2 - l2))
        except Exception as ex:
            errors["checksum"] = repr(ex)
        if world == 1:
            try:
                legs["small_calls"] = BL.small_call_leg(cpu)
            except Exception as ex:
                errors["small_calls"] = repr(ex)
            try:
                legs["one_shot_calls"] = BL.one_shot_leg(cpu, BL.scaled(64 << 20, args.scale))
            except Exception as ex:
                errors["one_shot_calls"] = repr(ex)
            try:
                legs["inflate_one_member"] = BL.single_member_leg(g, cpu, BL.scaled(256 << 20, args.scale))
            except Exception as ex:
                errors["inflate_one_member"] = repr(ex)
            try:
                legs["chunk_carry"] = BL.chunk_carry_leg(g, cpu, BL.scaled(512 << 20, args.scale))
            except Exception as ex:
                errors["chunk_carry"] = repr(ex)
        try:                                            # the host link with every rank copying at once, no kernels
            legs["link_probe"] = BL.link_probe(g)
        except Exception as ex:
            errors["link_probe"] = repr(ex)
        # one process driving every GPU of the box (zb200_multi_*) must give the single-GPU bytes: rank 0, others idle
        if world > 1:
            g.barrier()
            if rank == 0:
                try:
                    legs["multi_gpu_one_process"] = multi_gate(zb, ctx, world)
                except Exception as ex:
                    errors["multi_gpu_one_process"] = repr(ex)
            g.barrier()

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": headline_config(total, world),
                "roofline": head["roofline"], "cpu_baseline": head.get("cpu_baseline"), "e2e": head["e2e"],
                "gpu_launches": head["gpu_launches"],
                "clocks": clocks.summary(), "parity": head["parity"], "ratio": head["ratio"],
                "size_vs_reference": head.get("size_vs_reference"),
                "legs": {k: v for k, v in legs.items() if k != "deflate_l1"},
                "kernels": head["kernels"], "bench_wall_s": round(time.time() - t_start, 1)}
        if isinstance(traffic, dict) and traffic:
            line["traffic_probe"] = traffic
        if errors:
            line["errors"] = errors
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    bad = [k for k, v in legs.items() for gk, gv in (v.get("parity") or {}).items() if gv is False]
    if bad:
        sys.stderr.write("PARITY FAILURE in legs: %s\n" % sorted(set(bad)))
        sys.exit(3)


def multi_gate(zb, ctx, world):
    """zb200_multi_* over all GPUs of the box == the single-GPU bytes (64 MiB of mixed data, levels 1 and 6)."""
    L = zb.lib()
    n = 64 << 20
    h = L.zb200_host_alloc(n)
    BL.fill(h, n, "mixed", 0)
    cap = L.zb200_deflate_bound(n, BL.CHUNK, zb.FRAME_GZIP)
    o1, o2 = L.zb200_host_alloc(cap), L.zb200_host_alloc(cap)
    m = C.c_void_p()
    r = L.zb200_multi_create(None, 0, C.byref(m))
    if r != 0:
        raise zb.ZB200Error(r, "zb200_multi_create")
    out = {"devices": int(L.zb200_multi_count(m))}
    try:
        for level in (1, 6):
            l1, l2 = C.c_size_t(cap), C.c_size_t(cap)
            r = L.zb200_deflate_host(ctx.handle, C.c_void_p(h), n, BL.CHUNK, level, 0, zb.FRAME_GZIP, 1, C.c_void_p(o1), C.byref(l1), None, None)
            if r != 0:
                raise zb.ZB200Error(r, "zb200_deflate_host")
            r = L.zb200_multi_deflate_host(m, C.c_void_p(h), n, BL.CHUNK, level, 0, zb.FRAME_GZIP, 1, C.c_void_p(o2), C.byref(l2), None, None)
            if r != 0:
                raise zb.ZB200Error(r, "zb200_multi_deflate_host")
            out["deflate_l%d_equals_single_gpu" % level] = bool(l1.value == l2.value and C.string_at(o1, l1.value) == C.string_at(o2, l2.value))
        c1, a1, c2, a2 = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        L.zb200_checksum_host(ctx.handle, C.c_void_p(h), n, 3, 0, 1, C.byref(c1), C.byref(a1))
        L.zb200_multi_checksum_host(m, C.c_void_p(h), n, 3, 0, 1, C.byref(c2), C.byref(a2))
        out["checksum_equals_single_gpu"] = bool((c1.value, a1.value) == (c2.value, a2.value))
    finally:
        L.zb200_multi_destroy(m)
        for p in (h, o1, o2):
            L.zb200_host_free(C.c_void_p(p))
    out["parity"] = {k: v for k, v in out.items() if isinstance(v, bool)}
    return out


if __name__ == "__main__":
    main()
