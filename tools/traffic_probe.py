"""DRAM-traffic probe for bench.py: one bounded pass of every leg of the metric, meant to run under
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file t.csv python tools/traffic_probe.py side.json
The side file records, per leg, how many of the library's kernels were launched (zb200_launch_count) and the
input bytes of the pass; attribute() joins it with ncu's per-launch list: DRAM bytes (read + write) summed over a
leg's launches / the leg's input bytes.  Byte counts only — nothing here is a timing."""
import csv
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = ("dfl_", "inflate_", "ck_", "flush_candidates", "gather_segments", "gz_candidates", "tables_selftest")


def attribute(csv_path, side_path):
    side = json.load(open(side_path))
    rows = []
    with open(csv_path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    per = {}
    order = []
    for r in rd:
        name = r.get("Kernel Name", "")
        if not any(name.startswith(p) or ("zb::" + p) in name for p in OURS):
            continue
        kid = r.get("ID")
        if kid not in per:
            per[kid] = {"name": name.split("(")[0].replace("zb::", ""), "bytes": 0.0}
            order.append(kid)
        try:
            v = float(r.get("Metric Value", "0").replace(",", ""))
        except ValueError:
            continue
        unit = (r.get("Metric Unit") or "byte").lower()
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        per[kid]["bytes"] += v
    out, pos = {}, 0
    for leg in side["legs"]:
        k = leg["launches"]
        ids = order[pos:pos + k]
        pos += k
        by_kernel = {}
        for i in ids:
            by_kernel[per[i]["name"]] = by_kernel.get(per[i]["name"], 0) + per[i]["bytes"]
        tot = sum(by_kernel.values())
        if leg["name"].startswith("_"):
            continue
        out[leg["name"]] = {"dram_bytes": int(tot), "input_bytes": leg["input_bytes"], "algorithmic_bytes": leg["algorithmic_bytes"],
                            "dram_bytes_per_input_byte": round(tot / leg["input_bytes"], 4),
                            "dram_over_algorithmic": round(tot / leg["algorithmic_bytes"], 3),
                            "by_kernel": {k2: int(v) for k2, v in sorted(by_kernel.items(), key=lambda kv: -kv[1])},
                            "launches_profiled": len(ids)}
    if pos != len(order):
        out["note"] = "%d profiled launches of the library, %d attributed" % (len(order), pos)
    return out


def main(side_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import bench_legs as BL
    import zlib_wasm_b200 as zb
    L = zb.lib()
    ctx = zb.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    legs = []
    n = 512 << 20                                       # one sub-batch of the deflate pipeline
    host = L.zb200_host_alloc(n)
    cap = L.zb200_deflate_bound(n, BL.CHUNK, zb.FRAME_GZIP_MEMBERS)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_end = torch.zeros(n // BL.CHUNK, dtype=torch.int64, device="cuda")
    h_view = torch.frombuffer((C.c_uint8 * n).from_address(host), dtype=torch.uint8)
    for gen in ("markov", "mixed"):
        BL.fill(host, n, gen, 0)
        d_in.copy_(h_view)
        torch.cuda.synchronize()
        for name, (kind, _, level, strategy) in BL.DEFLATE_LEGS.items():
            if kind != gen:
                continue
            l0 = L.zb200_launch_count()
            r = L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), n, BL.CHUNK, level, strategy, zb.FRAME_RAW, 1, d_out.data_ptr(), cap,
                                    None, d_tot.data_ptr(), sp)
            assert r == 0, zb.last_error()
            torch.cuda.synchronize()
            legs.append({"name": name, "launches": int(L.zb200_launch_count() - l0), "input_bytes": n, "algorithmic_bytes": n + int(d_tot.item())})
        if gen == "markov":                             # inflate: members of 256 KiB made from the same text
            l0 = L.zb200_launch_count()
            r = L.zb200_deflate_dev(ctx.handle, d_in.data_ptr(), n, BL.CHUNK, 6, 0, zb.FRAME_GZIP_MEMBERS, 1, d_out.data_ptr(), cap,
                                    d_end.data_ptr(), d_tot.data_ptr(), sp)
            assert r == 0, zb.last_error()
            torch.cuda.synchronize()
            # making the members is not a leg: its launches go to a throw-away entry
            legs.append({"name": "_setup_members", "launches": int(L.zb200_launch_count() - l0), "input_bytes": n, "algorithmic_bytes": n})
            ends = d_end.cpu().tolist()
            members, prev = [], 0
            for i, e in enumerate(ends):
                members.append(zb.Member(prev, e - prev, i * BL.CHUNK, BL.CHUNK, 0, 0, 0))
                prev = e
            arr = (zb.Member * len(members))(*members)
            d_members = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
            d_res = torch.zeros(len(members) * C.sizeof(zb.MemberResult), dtype=torch.uint8, device="cuda")
            d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
            l0 = L.zb200_launch_count()
            r = L.zb200_inflate_dev(ctx.handle, d_out.data_ptr(), d_back.data_ptr(), d_members.data_ptr(), len(members), zb.WRAP_GZIP, 1,
                                    d_res.data_ptr(), sp)
            assert r == 0, zb.last_error()
            torch.cuda.synchronize()
            legs.append({"name": "inflate", "launches": int(L.zb200_launch_count() - l0), "input_bytes": n, "algorithmic_bytes": n + prev})
            del d_back
    l0 = L.zb200_launch_count()
    d_out2 = torch.zeros(2, dtype=torch.int32, device="cuda")
    r = L.zb200_checksum_dev(ctx.handle, d_in.data_ptr(), n, 3, 0, 1, d_out2.data_ptr(), sp)
    assert r == 0
    torch.cuda.synchronize()
    legs.append({"name": "checksum", "launches": int(L.zb200_launch_count() - l0), "input_bytes": n, "algorithmic_bytes": n})
    json.dump({"legs": legs}, open(side_path, "w"))


if __name__ == "__main__":
    main(sys.argv[1])
