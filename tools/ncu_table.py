import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = [("ms", "gpu__time_duration.sum", 1e-6), ("issue active %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1), ("warp inst (M)", "smsp__inst_executed.sum", 1e-6),
        ("lanes/inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1), ("warps active/SM", "sm__warps_active.avg.per_cycle_active", 1),
        ("regs", "launch__registers_per_thread", 1), ("occ limit: barriers", "launch__occupancy_limit_barriers", 1), ("occ limit: smem", "launch__occupancy_limit_shared_mem", 1),
        ("occ limit: regs", "launch__occupancy_limit_registers", 1), ("occ limit: warps", "launch__occupancy_limit_warps", 1),
        ("DRAM read MB", "dram__bytes_read.sum", 1e-6), ("DRAM write MB", "dram__bytes_write.sum", 1e-6), ("DRAM % of peak", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("L2 hit %", "lts__t_sector_hit_rate.pct", 1), ("smem bank conflicts (M)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1e-6),
        ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1)]
units = rows[1]
names = []
cols = {}
for r in rows[2:]:
    k = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("zb::", "")
    inst = k
    i = 2
    while inst in cols:
        inst = "%s #%d" % (k, i); i += 1
    names.append(inst)
    cols[inst] = r
sel = names if len(sys.argv) < 3 else [n for n in names if any(s in n for s in sys.argv[2].split(","))]
print("%-26s" % "metric" + "".join("%22s" % n[:21] for n in sel))
for label, m, sc in want:
    if m not in hdr:
        continue
    j = hdr.index(m)
    vals = []
    for n in sel:
        v = cols[n][j].replace(",", "")
        try:
            x = float(v)
            u = units[j]
            if m == "gpu__time_duration.sum":
                x *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
            elif sc != 1:
                mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
                x = x * mult * sc
            vals.append("%22.3f" % x)
        except ValueError:
            vals.append("%22s" % v[:21])
    print("%-26s" % label + "".join(vals))
