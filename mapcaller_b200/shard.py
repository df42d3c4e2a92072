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


# ---- genome tiles of mc_profile_reduce_scatter (csrc/mc_ctx.cu: profile_reduce) ----------------------------------------------
TILE_ALIGN = 25600   # columns: a multiple of the 1024-column prefix blocks and of the 100-column variant-scan blocks


def tile_size(genome: int, world: int) -> int:
    """Columns per rank: equal tiles of whole TILE_ALIGN units that cover the genome."""
    per = (genome + world - 1) // world
    return (per + TILE_ALIGN - 1) // TILE_ALIGN * TILE_ALIGN


def tile_bounds(genome: int, world: int, rank: int):
    """[begin, end) columns whose counters `rank` holds after the reduce-scatter (possibly empty for the last ranks of a tiny genome)."""
    t = tile_size(genome, world)
    return min(genome, rank * t), min(genome, (rank + 1) * t)


def prefix_with_carry(tile_diff: np.ndarray, totals_of_all_ranks, rank: int) -> np.ndarray:
    """Coverage of the tile's columns from its difference-array entries: the running sum inside the tile plus the sums of all
    entries before it = the tile totals of the earlier ranks (what profile_prefix plants in front of the tile)."""
    carry = sum(totals_of_all_ranks[:rank])
    return np.cumsum(tile_diff, axis=0) + carry


def run_carry(last_seen_of_all_ranks, rank: int) -> int:
    """Last non-gap (non-dup) column before the tile: the maximum over the earlier ranks, -1 if none (mc_variant_scan)."""
    return max([-1] + [int(x) for x in last_seen_of_all_ranks[:rank]])
