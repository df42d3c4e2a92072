"""Host-side sharding of a read library over the GPUs of one box.

Reads are independent units given the index, so the library is cut into contiguous, chunk-aligned shards in file order:
rank r maps chunks [r*C/N, (r+1)*C/N) of the 200-read chunk grid (reference chunk protocol, src/GetData.cpp:85-99).
Everything additive is reduced afterwards (mc_profile_allreduce, NCCL)."""
from __future__ import annotations

import numpy as np

CHUNK_READS = 200


def shard_bounds(n_reads: int, world: int, rank: int, paired: bool = True):
    """[begin, end) read indices of `rank`'s shard; shards are contiguous, cover the library and start on chunk borders."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    n_chunks = (n_reads + CHUNK_READS - 1) // CHUNK_READS
    b = (n_chunks * rank // world) * CHUNK_READS
    e = min(n_reads, (n_chunks * (rank + 1) // world) * CHUNK_READS)
    if paired and (b & 1 or (e & 1 and e != n_reads)):
        raise AssertionError("chunk grid keeps mates together")
    return b, max(b, e)


def take_shard(seq: np.ndarray, off: np.ndarray, world: int, rank: int, paired: bool = True):
    """(seq, off) view of the shard with offsets rebased to 0."""
    b, e = shard_bounds(len(off) - 1, world, rank, paired)
    return seq[off[b]:off[e]], off[b:e + 1] - off[b]


def merge_totals(parts):
    """Sum of the per-shard totals (iTotalReadNum, iTotalMappingNum, iTotalPairedNum, TotalPairedDistance, ReadLengthSum)."""
    keys = ("total_reads", "total_mapped", "total_paired", "total_distance", "read_length_sum")
    out = {k: int(sum(p[k] for p in parts)) for k in keys}
    out["avg_dist"] = int(1.0 * out["total_distance"] / out["total_paired"] + 0.5) if out["total_paired"] > 1000 else 1000
    return out
