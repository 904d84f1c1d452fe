"""Levels 1-3: compressed size next to the reference's, per data generator (the streams are not the reference's byte for byte:
the tolerance north_star sets is 3 %).  One-shot calls of 1 MiB (one run) and 8 MiB (256 KiB chunks on both sides)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refz  # noqa: E402
import zlib_wasm_b200 as zb  # noqa: E402

z = refz.ZlibBinding(zb.LIB_PATH, "")
ref = refz.ref()
names = {refz.GEN_TEXT: "text", refz.GEN_MARKOV: "markov", refz.GEN_RANDOM: "random", refz.GEN_MIXED: "mixed", refz.GEN_BYTES: "bytes"}
worst = 0.0
for level in (1, 2, 3):
    for kind, name in names.items():
        row = []
        for n, chunk in ((1 << 20, 0), (8 << 20, 262144)):
            d = refz.gen(n, kind, seed=7 * kind + 1)
            a = z.deflate_stream(d, level, 0, refz.WRAP_RAW, chunk)
            b = ref.deflate_stream(d, level, 0, refz.WRAP_RAW, chunk)
            r = len(a) / len(b)
            worst = max(worst, r)
            row.append("%d MiB %.4f" % (n >> 20, r))
        print("level %d %-7s %s" % (level, name, "   ".join(row)), flush=True)
print("worst %.4f" % worst)
