"""Generate tests/golden/golden.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/libzref.so):
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the committed JSON.

Contents
  puff_vectors  the malformed raw-deflate snippets of contrib/puff/Makefile:19-38
                with the reference inflate()'s (ret, msg) for each
  zeros_raw     contrib/puff/zeros.raw (2517 B known-answer stream) + its output
                length / CRC-32 / Adler-32 as decoded by the reference
  streams       reference deflate output (hex) for small seeded inputs at
                level x strategy x wrap x chunking, + CRC-32/Adler-32 of the input
  checksums     crc32_z / adler32_z / combine known answers at the edge sizes
                SURVEY.md §7 lists (0,1,15,16,39..48,5551..5553, unaligned)
"""
import base64
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refz  # noqa: E402

REF = "/root/reference"


def main():
    r = refz.ref()
    g = {"reference_version": r.version.decode(), "seed": refz.SEED}

    # --- puff malformed vectors ------------------------------------------
    vec = []
    mk = open(os.path.join(REF, "contrib/puff/Makefile")).read()
    for m in re.finditer(r"@echo '([0-9a-f ]+)' \| xxd -r -p \| puft( -f)?.*?(?:-eq (\d+))?$", mk, re.M):
        raw = bytes.fromhex(m.group(1).replace(" ", ""))
        ret, msg, out, tin = r.inflate_all(raw, refz.WRAP_RAW, cap=4096)
        vec.append({"hex": raw.hex(), "puff_exit": int(m.group(3)) if m.group(3) else 0,
                    "ret": ret, "msg": msg, "out_hex": out.hex(), "total_in": tin})
    g["puff_vectors"] = vec

    # --- zeros.raw --------------------------------------------------------
    z = open(os.path.join(REF, "contrib/puff/zeros.raw"), "rb").read()
    ret, msg, out, tin = r.inflate_all(z, refz.WRAP_RAW, cap=2000000)
    assert ret == refz.Z_STREAM_END and set(out) == {0}
    g["zeros_raw"] = {"b64": base64.b64encode(z).decode(), "out_len": len(out),
                      "crc32": r.crc32_z(0, out, len(out)), "adler32": r.adler32_z(1, out, len(out)),
                      "total_in": tin}

    # --- reference deflate streams ---------------------------------------
    streams = []
    inputs = [("text", refz.GEN_TEXT, 3000), ("markov", refz.GEN_MARKOV, 70000), ("random", refz.GEN_RANDOM, 1500),
              ("mixed", refz.GEN_MIXED, 140000), ("empty", refz.GEN_TEXT, 0), ("one", refz.GEN_TEXT, 1)]
    for name, kind, n in inputs:
        data = refz.gen(n, kind)
        for level in (1, 3, 4, 6, 9):
            for strategy in (0, 1, 2, 3, 4):
                if strategy and level not in (1, 6):
                    continue
                for wrap in (0, 1, 2):
                    for chunk in (0, 32768):
                        if chunk and (n <= chunk or wrap == 0 and level != 6):
                            continue
                        s = r.deflate_stream(data, level, strategy, wrap, chunk)
                        e = {"input": name, "kind": kind, "n": n, "level": level, "strategy": strategy,
                             "wrap": wrap, "chunk": chunk, "len": len(s),
                             "sha256": hashlib.sha256(s).hexdigest()}
                        if len(s) <= 2500:
                            e["hex"] = s.hex()
                        streams.append(e)
        # fingerprint of the generator output so a drifting generator is caught
        streams.append({"input": name, "kind": kind, "n": n, "data_sha256": hashlib.sha256(data).hexdigest(),
                        "crc32": r.crc32_z(0, data, n), "adler32": r.adler32_z(1, data, n)})
    g["streams"] = streams

    # --- checksum edge sizes ---------------------------------------------
    big = refz.gen(70000, refz.GEN_BYTES)
    cs = []
    for n in [0, 1, 2, 3, 15, 16, 17, 31, 32, 39, 40, 41, 46, 47, 48, 63, 64, 65, 255, 256, 4095, 4096,
              5551, 5552, 5553, 11104, 11105, 65535, 65536, 65537, 69999]:
        for off in (0, 1, 3, 7):
            if off + n > len(big):
                continue
            d = big[off:off + n]
            cs.append({"off": off, "n": n, "crc32": r.crc32_z(0, d, n), "adler32": r.adler32_z(1, d, n),
                       "crc32_seeded": r.crc32_z(0xdeadbeef, d, n), "adler32_seeded": r.adler32_z(0x12345678 % (65521 << 16) | 5, d, n)})
    comb = []
    for l2 in [0, 1, 2, 255, 256, 65535, 65536, 262144, (1 << 32) - 1, 1 << 32, (1 << 33) + 12345]:
        comb.append({"len2": l2, "crc": r.crc32_combine(0x12345678, 0x9abcdef0, l2),
                     "gen": r.crc32_combine_gen(l2), "adler": r.adler32_combine(0x00c8012d, 0x11e60398, l2)})
    g["checksums"] = {"data_kind": refz.GEN_BYTES, "data_n": 70000, "cases": cs, "combine": comb,
                      "null": {"crc32": r.crc32(0, None, 0), "adler32": r.adler32(0, None, 0)}}

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    print("wrote golden.json:", len(vec), "puff vectors,", len(streams), "stream entries,", len(cs), "checksum cases")


if __name__ == "__main__":
    main()
