"""Per-source-line view of an ncu report: joins the SASS page of `ncu --set full
--import-source on` (executed instructions, stall samples per instruction) with the
line table of the cubin (nvdisasm -g), because the CUDA-source page of the CLI
carries no metrics.  Usage: python tools/ncu_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTR [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout
# walk the functions; record (offset -> (file, line, inline chain)) for the requested kernel
line_of, cur, in_k = {}, None, False
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        in_k = kern in m.group(1)
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and in_k:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
stall_cols = [(n, i) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
base = None
agg = collections.OrderedDict()
stalls = collections.defaultdict(lambda: collections.Counter())
tot_i = tot_s = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[0].strip():
        continue
    try:
        a = int(r[ix["Address"]], 16) if not r[ix["Address"]].isdigit() else int(r[ix["Address"]])
    except ValueError:
        continue
    if base is None:
        base = a
    off = a - base
    info = line_of.get(off, (None, ""))[0]
    key = (info[0], info[1]) if info else ("?", 0)
    n_i = int(r[ix["Instructions Executed"]] or 0)
    n_s = int(r[ix["# Samples"]] or 0)
    n_t = int(r[ix["Thread Instructions Executed"]] or 0)
    g = agg.setdefault(key, [0, 0, 0, 0])
    g[0] += n_i; g[1] += n_s; g[2] += n_t; g[3] += 1
    for n, i in stall_cols:
        v = int(r[i] or 0)
        if v:
            stalls[key][n[6:]] += v
    tot_i += n_i; tot_s += n_s
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
print("%-28s %6s %12s %7s %9s %7s %6s" % ("file:line", "sass", "warp_inst", "inst%", "samples", "smp%", "lanes"))
for key, g in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    why = " ".join("%s=%d%%" % (n, 100 * v // max(g[1], 1)) for n, v in stalls[key].most_common(3))
    print("%-28s %6d %12d %6.2f%% %9d %6.2f%% %6.1f  %s" % ("%s:%d" % key, g[3], g[0], 100.0 * g[0] / max(tot_i, 1), g[1], 100.0 * g[1] / max(tot_s, 1), g[2] / max(g[0], 1), why))
