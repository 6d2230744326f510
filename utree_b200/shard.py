"""Read sharding for one-process-per-GPU runs (bench.py under torchrun): reads
are independent units, the CTR is replicated, so a rank just takes a contiguous
range of records and rank 0 concatenates the per-rank outputs in rank order --
no data-path collective."""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int):
    """Records [lo, hi) owned by `rank`: contiguous, balanced to within one."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def split_fasta(data: bytes, rank: int, world: int):
    """Byte range of a 2-line-per-record FASTA owned by `rank` (cuts on record boundaries)."""
    starts = [0]
    pos, line = 0, 0
    while True:
        nl = data.find(b"\n", pos)
        if nl < 0:
            break
        pos = nl + 1
        line += 1
        if line % 2 == 0 and pos < len(data):
            starts.append(pos)
    n = len(starts)
    lo, hi = shard_range(n, rank, world)
    starts.append(len(data))
    return starts[lo], starts[hi] if hi < n else len(data)


def merge_outputs(parts):
    """Ordered merge: rank order == input order."""
    return b"".join(parts)
