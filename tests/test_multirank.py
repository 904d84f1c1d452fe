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


CARRY_WORKER = r'''
import ctypes as C, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
import refz, test_emul
from zlib_wasm_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
CH, W = 65536, 32768
total = 7 * CH + 4321
lo, hi = shard.shard_range(total, rank, world, align=CH)
whole = refz.gen(total, refz.GEN_TEXT, seed=9)
# a rank takes its chunks PLUS the 32 KiB before its range (zb200_multi_deflate_host copies them along; a torchrun rank reads them
# from the shared input): the first chunk of rank > 0 is compressed behind its neighbour's tail
L = test_emul._build("def_emul")
L.emul_deflate_chunk_dict.restype = C.c_long
L.emul_deflate_chunk_dict.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
mine = []
for pos in range(lo, hi, CH):
    hist, piece = whole[max(0, pos - W):pos], whole[pos:min(pos + CH, hi)]
    last = pos + CH >= total
    joined = hist + piece
    cap = len(joined) + len(joined) // 8 + 1024
    out, st = C.create_string_buffer(cap), (C.c_uint32 * 2)()
    r = L.emul_deflate_chunk_dict(joined, len(joined), len(hist), 6, 0, 1 if last else 0, out, cap, st)
    assert r >= 0
    mine.append(out.raw[:r])
got = [None] * world
dist.all_gather_object(got, b"".join(mine))
stream = b"".join(got)                                  # streams concatenated in rank order: the exchange step
if rank == 0 and refz.have_ref():
    ref = refz.ref()
    want = b"".join(ref.deflate_stream(whole[p:p + CH], 6, 0, refz.WRAP_RAW, 0, dictionary=whole[max(0, p - W):p] or None,
                                       last_flush=refz.Z_FINISH if p + CH >= total else refz.Z_SYNC_FLUSH) for p in range(0, total, CH))
    assert stream == want, (len(stream), len(want))
    err, msg, back, used = ref.inflate_all(stream, refz.WRAP_RAW, cap=total + 16)
    assert err == 1 and back == whole
else:
    import zlib
    assert zlib.decompress(stream, -15) == whole
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_gloo_ranks_carried_history(tmp_path):
    """Chunks sharded over two ranks with history carried across the rank boundary (ZB200_CHUNK_CARRY in zb200_multi_* / a
    torchrun job): the host replay of the device cores stands in for the kernels; the concatenated stream is the
    reference's per-chunk dictionary construction and decodes to the input."""
    script = tmp_path / "carry_worker.py"
    script.write_text(CARRY_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), refz.ROOT, port, str(r), "2"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
