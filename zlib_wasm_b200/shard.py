"""Host-side multi-GPU logic: the path shards with no data-path
collective.  Contiguous ranges of units (bytes / chunks / members) per rank;
per-rank partial checksums are merged left to right with the reference's combine
algebra (crc32.c:1021 crc32_combine, adler32.c:133 adler32_combine), exposed by
the library as pure host functions.  torch.distributed is only the plumbing for the
gather (NCCL on GPUs, gloo in the CPU tests)."""
from . import lib


def shard_range(total, rank, world, align=1):
    """Contiguous [start, end) of `total` units for `rank`; boundaries are multiples
    of `align` (e.g. the deflate chunk size) except the very end."""
    units = (total + align - 1) // align
    per, extra = divmod(units, world)
    first = rank * per + min(rank, extra)
    count = per + (1 if rank < extra else 0)
    return min(first * align, total), min((first + count) * align, total)


def combine_checksums(parts):
    """parts: [(crc32, adler32, nbytes), ...] in stream order -> (crc32, adler32, nbytes)."""
    L = lib()
    crc, adler, n = 0, 1, 0
    for c, a, k in parts:
        crc = L.zb200_crc32_combine(crc, c, k)
        adler = L.zb200_adler32_combine(adler, a, k)
        n += k
    return crc, adler, n


def gather_checksums(dist, crc, adler, nbytes, device="cpu"):
    """All-gather every rank's (crc, adler, nbytes) and combine them in rank order."""
    import torch
    world = dist.get_world_size() if dist.is_initialized() else 1
    mine = torch.tensor([crc, adler, nbytes], dtype=torch.int64, device=device)
    if world == 1:
        return combine_checksums([tuple(int(x) for x in mine.tolist())])
    got = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    return combine_checksums([tuple(int(x) for x in g.tolist()) for g in got])
