"""N>1 host logic on CPU: two gloo ranks each checksum their own shard (with the
oracle standing in for the GPU kernel — this test is about sharding + combine, not
about the kernels) and the gathered, combined result must equal the checksum of
the whole buffer.  Also covers chunk-aligned shard boundaries."""
import os
import subprocess
import sys

import refz
from zlib_wasm_b200 import shard

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
import refz
from zlib_wasm_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
total = 5 * 262144 + 12345
lo, hi = shard.shard_range(total, rank, world, align=262144)
assert lo % 262144 == 0
blocks = 65536
data = refz.gen(total, refz.GEN_BYTES)[lo:hi]       # every rank can regenerate any range: generators are block-parallel
o = refz.oracle()
crc, adler, n = shard.gather_checksums(dist, o.crc32(data), o.adler32(data), len(data))
whole = refz.gen(total, refz.GEN_BYTES)
assert n == total and (crc, adler) == (o.crc32(whole), o.adler32(whole)), (rank, crc, adler)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_shard_range_properties():
    for total in (0, 1, 262143, 262144, 10 * 262144 + 5):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(total, r, world, 262144) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(s[0] % 262144 == 0 for s in spans if s[1] > s[0])


def test_combine_matches_whole_buffer():
    o = refz.oracle()
    d = refz.gen(700001, refz.GEN_MIXED)
    cuts = [0, 100, 262144, 524288, 700001]
    parts = [(o.crc32(d[a:b]), o.adler32(d[a:b]), b - a) for a, b in zip(cuts, cuts[1:])]
    assert shard.combine_checksums(parts) == (o.crc32(d), o.adler32(d), len(d))


def test_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), refz.ROOT, port, str(r), "2"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
