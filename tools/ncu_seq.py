"""Per-call kernel time sequences (us) from an ncu launch-list csv."""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
cur = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1)
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("zb::", "").strip()
    if n.startswith(("dfl_", "inflate", "ck_")):
        cur.append("%s=%.0f" % (n.replace("dfl_", "").replace("_kernel", ""), v))
    if n in ("dfl_frame_kernel", "inflate_verify_kernel"):
        print(" ".join(cur)); cur = []
