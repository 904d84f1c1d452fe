"""Generate tests/golden/carry_golden.json from the UNMODIFIED reference (needs oracle/_ref/libzref.so):
    python tests/golden/make_carry_golden.py
Carried-history chunk streams (zb200.h ZB200_CHUNK_CARRY): per chunk the reference's deflateInit2(raw) +
deflateSetDictionary(the 32 KiB before the chunk) + deflate(chunk, Z_SYNC_FLUSH; the last one Z_FINISH), laid end to end
(deflate.c:550-632,1211-1218).  Stored per case: the stream's length and SHA-256, and for the smallest case the bytes."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refz  # noqa: E402

CASES = [  # kind, n, chunk, level, strategy
    (refz.GEN_TEXT, 3000, 1000, 6, 0), (refz.GEN_TEXT, 300000, 65536, 6, 0), (refz.GEN_TEXT, 300000, 65536, 9, 0),
    (refz.GEN_MARKOV, 150000, 100000, 4, 0), (refz.GEN_MARKOV, 150000, 20000, 6, 1), (refz.GEN_MIXED, 200000, 20000, 6, 0),
    (refz.GEN_MIXED, 200000, 32768, 8, 0), (refz.GEN_BYTES, 120000, 40000, 6, 3), (refz.GEN_TEXT, 120000, 40000, 6, 2),
    (refz.GEN_RANDOM, 70000, 30000, 6, 0), (refz.GEN_TEXT, 70000, 262144, 6, 0), (refz.GEN_MARKOV, 262144 + 5, 262144, 5, 0),
]


def carried(ref, d, chunk, level, strategy):
    n = len(d)
    nch = max(1, (n + chunk - 1) // chunk)
    out = []
    for c in range(nch):
        pos = c * chunk
        hist = d[max(0, pos - 32768):pos]
        out.append(ref.deflate_stream(d[pos:pos + chunk], level, strategy, refz.WRAP_RAW, 0, dictionary=hist if hist else None,
                                      last_flush=refz.Z_FINISH if c == nch - 1 else refz.Z_SYNC_FLUSH))
    return out


def main():
    ref = refz.ref()
    g = {"reference_version": ref.version.decode(), "cases": []}
    for kind, n, chunk, level, strategy in CASES:
        d = refz.gen(n, kind, seed=4000 + kind)
        parts = carried(ref, d, chunk, level, strategy)
        s = b"".join(parts)
        e = {"kind": kind, "n": n, "seed": 4000 + kind, "chunk": chunk, "level": level, "strategy": strategy, "len": len(s),
             "chunk_lens": [len(p) for p in parts], "sha256": hashlib.sha256(s).hexdigest()}
        if n <= 3000:
            e["hex"] = s.hex()
        g["cases"].append(e)
    with open(os.path.join(HERE, "carry_golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", len(g["cases"]), "cases")


if __name__ == "__main__":
    main()
