"""The reference's own examples/zpipe.c (16 KiB slices through deflate() / inflate()), linked against libzb200.so and against
the reference: wall time of `zpipe < file` and `zpipe -d < file.z`, and whether the two produce the same bytes."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d = refz.gen(mib << 20, refz.GEN_TEXT, seed=0x9E37)
exes = {"b200": os.path.join(ROOT, "tests", "_bin", "zpipe_b200"), "reference": os.path.join(ROOT, "oracle", "_ref", "zpipe")}
outs = {}
for name, exe in exes.items():
    if not os.path.exists(exe):
        continue
    subprocess.run([exe], input=d[:1 << 20], capture_output=True)      # warm-up (device buffers, page cache)
    t0 = time.perf_counter(); c = subprocess.run([exe], input=d, capture_output=True); t1 = time.perf_counter()
    u = subprocess.run([exe, "-d"], input=c.stdout, capture_output=True); t2 = time.perf_counter()
    outs[name] = c.stdout
    print("%-9s zpipe %d MiB: compress %8.1f ms -> %d bytes (rc %d), decompress %8.1f ms (rc %d, %s)" %
          (name, mib, (t1 - t0) * 1e3, len(c.stdout), c.returncode, (t2 - t1) * 1e3, u.returncode, "ok" if u.stdout == d else "DIFFERENT"), flush=True)
if len(outs) == 2:
    print("same compressed bytes:", outs["b200"] == outs["reference"])
