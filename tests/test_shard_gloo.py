"""N > 1 host logic on CPU: two gloo ranks shard one library, map their shards with the CPU oracle (as separate
libraries, which is what the round-1 multi-GPU path does) and reduce the additive results the way
mc_profile_allreduce does with NCCL; the reduced totals / profile must equal the sum of the per-shard oracle runs and
the shards must tile the library."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import parity_util as pu
    from mapcaller_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = pu.make_case(seed=51, n_pairs=1500, genome_len=50000)
    ix = pu.build_index(case)
    seq, off = shard.take_shard(case["seq"], case["off"], world, rank)
    mine = pu.oracle_results(dict(case, seq=seq, off=off), ix, want_reads=False)
    # what NCCL does on the GPU: sum the additive arrays, gather the records
    prof = torch.from_numpy(mine["profile"].astype(np.int64))
    dist.all_reduce(prof)
    c = mine["counters"]
    t = torch.tensor([c["reads"], c["mapped"], c["paired"], c["dist_sum"], c["len_sum"]], dtype=torch.int64)
    dist.all_reduce(t)
    recs = [None] * world
    dist.all_gather_object(recs, (mine["ins"], mine["dele"], mine["bp"], shard.shard_bounds(len(case["off"]) - 1, world, rank)))
    if rank == 0:
        q.put((prof.numpy(), t.tolist(), recs, mine["profile"], len(case["off"]) - 1))
    dist.destroy_process_group()


def test_two_rank_shard_and_reduce(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    prof, tot, recs, prof0, n_reads = q.get(timeout=300)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile the library on chunk borders
    (b0, e0), (b1, e1) = recs[0][3], recs[1][3]
    assert b0 == 0 and e0 == b1 and e1 == n_reads and e0 % 200 == 0
    assert tot[0] == n_reads and tot[1] > 0.9 * n_reads
    # the reduced profile is the sum of the shard profiles (rank 0's share is a part of it)
    assert (prof >= prof0).all() and prof[:, :4].sum() > prof0[:, :4].sum()


def test_shard_bounds_cover_any_library():
    from mapcaller_b200 import shard
    for n in (0, 2, 198, 200, 202, 4000, 123456):
        for w in (1, 2, 3, 4, 8):
            cuts = [shard.shard_bounds(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0] and a[1] % 200 == 0
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 2, 2)


def _tile_worker(rank, world, port, q):
    """The read-out protocol of mc_profile_reduce_scatter with numpy arrays over gloo: every rank holds partial difference arrays
    of the whole genome; after the reduce each keeps its tile; coverage prefixes and gap-run carriers of the earlier tiles come
    from small all-gathers (csrc/mc_ctx.cu: profile_reduce, profile_prefix, mc_variant_scan)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mapcaller_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = 90_000
    rng = np.random.default_rng(100 + rank)
    diff = np.zeros((G + 1, 2), dtype=np.int64)                      # two lanes of a difference array: +1 at a start, -1 at the end
    starts = rng.integers(30_000 * (rank == 1), G - 300, size=400)   # rank 1's reads leave the first 30 k columns empty
    for s in starts:
        diff[s, rank % 2] += 1; diff[s + int(rng.integers(50, 300)), rank % 2] -= 1
    total = torch.from_numpy(diff.copy()); dist.all_reduce(total)    # (gloo has no reduce-scatter: reduce, then keep the own tile)
    b, e = shard.tile_bounds(G, world, rank)
    tile = total.numpy()[b:e]
    tot = [None] * world; dist.all_gather_object(tot, tile.sum(axis=0))
    cov = shard.prefix_with_carry(tile, tot, rank)
    covered = np.nonzero(cov.sum(axis=1) > 0)[0]
    last = [None] * world; dist.all_gather_object(last, int(b + covered[-1]) if len(covered) else -1)
    carry = shard.run_carry(last, rank)
    out = [None] * world
    dist.all_gather_object(out, (b, e, cov, carry))
    if rank == 0:
        q.put((total.numpy(), out))
    dist.destroy_process_group()


def test_reduce_scatter_read_out_protocol():
    import torch.multiprocessing as mp
    from mapcaller_b200 import shard
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_tile_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    total, out = q.get(timeout=300)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    G = 90_000
    whole = np.cumsum(total[:G], axis=0)
    assert out[0][:2] == (0, 51_200) and out[1][:2] == (51_200, G)
    for b, e, cov, carry in out:
        assert np.array_equal(cov, whole[b:e])                                       # tile prefix + carry == prefix over the whole genome
        seen = np.nonzero(whole[:b].sum(axis=1) > 0)[0]
        assert carry == (int(seen[-1]) if len(seen) else -1)                         # last covered column before the tile
    # tiles are equal, aligned, and cover any genome for any rank count
    for g in (1, 99, 25_600, 25_601, 248_956_422, 3_088_269_832):
        for w in (1, 2, 3, 8, 16):
            t = shard.tile_size(g, w)
            assert t % shard.TILE_ALIGN == 0 and t * w >= g and t * w <= g + w * shard.TILE_ALIGN
            cuts = [shard.tile_bounds(g, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == g and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
