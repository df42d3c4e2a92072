"""Two contexts on two GPUs of one box (one process per GPU, torchrun, NCCL over NVLink).

* ordered exchange (the default once mc_comm_init was called): the ranks map consecutive shards of ONE library; the avgDist
  trajectory, the dedup gate and the discordant-pair state are exchanged in file order, so every per-read record, the
  EstiDistance of every chunk, the totals and - after mc_profile_allreduce - the profile equal the oracle's single run over
  the whole library;
* independent shards (mc_params.reserved[2] = 1): each rank maps its shard as a library of its own and
  mc_profile_allreduce leaves the sum on both ranks; checked against the oracle run on the same two shards."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import parity_util as pu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        return len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
    except OSError:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_gpu_shards_reduce_to_the_sum(built, tmp_path):
    code = textwrap.dedent("""
        import os, sys, pickle
        import numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        import parity_util as pu
        from mapcaller_b200 import api, shard
        rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
        torch.cuda.set_device(rank); dist.init_process_group('nccl')
        case = pu.make_case(seed=61, n_pairs=8000, genome_len=120000, contigs=2, sv=2.0)
        ix = pu.build_index(case)
        seq, off = shard.take_shard(case['seq'], case['off'], world, rank)
        ctx = api.Context(ix, paired=1, device=rank, shard_rank=rank, shard_count=world, reserved=(0, 0, 1, 0, 0))
        uid = [api.Context.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        ctx.map_batch(seq, off)
        ctx.profile_allreduce()
        ins, dele = ctx.indels()
        out = dict(profile=ctx.profile_columns(), totals=ctx.totals(), ins=ins, dele=dele, bp=ctx.breakpoints(), inv=sorted(ctx.sites(0)), tnl=sorted(ctx.sites(1)))
        pickle.dump(out, open(%r + '/r%%d.pkl' %% rank, 'wb'))
        dist.destroy_process_group()
    """) % (ROOT, os.path.join(ROOT, "tests"), str(tmp_path))
    script = tmp_path / "w.py"
    script.write_text(code)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)])
    import pickle
    from collections import Counter
    from mapcaller_b200 import shard
    got = [pickle.load(open(str(tmp_path / ("r%d.pkl" % r)), "rb")) for r in range(2)]
    case = pu.make_case(seed=61, n_pairs=8000, genome_len=120000, contigs=2, sv=2.0)
    ix = pu.build_index(case)
    parts = []
    for r in range(2):
        seq, off = shard.take_shard(case["seq"], case["off"], 2, r)
        parts.append(pu.oracle_results(dict(case, seq=seq, off=off), ix, want_reads=False))
    want = parts[0]["profile"].astype(np.int64) + parts[1]["profile"].astype(np.int64)
    want[:, :5] = np.minimum(want[:, :5], 4095); want[:, 5] = np.minimum(want[:, 5], 5)
    for g in got:                                           # both ranks hold the reduced profile
        assert np.array_equal(g["profile"], want)
        assert g["totals"]["total_reads"] == 16000
        assert g["totals"]["total_paired"] == parts[0]["counters"]["paired"] + parts[1]["counters"]["paired"]
        for key in ("ins", "dele"):
            c = Counter()
            for p in parts:
                for pos, s, n in p[key]:
                    c[(pos, s)] += n
            assert sorted((k[0], k[1], v) for k, v in c.items()) == sorted(g[key])
        assert sorted(g["inv"]) == sorted(parts[0]["inv"] + parts[1]["inv"])
        assert sorted(g["tnl"]) == sorted(parts[0]["tnl"] + parts[1]["tnl"])


CASE = dict(seed=62, n_pairs=20000, genome_len=120000, contigs=2, sv=2.0, n_dup=10, frag_mean=380, frag_sd=60)


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("super_batches,paired", [(1, 1), (3, 1), (2, 0)])
def test_two_gpus_one_library_equals_single_reference_run(built, tmp_path, super_batches, paired):
    code = textwrap.dedent("""
        import os, sys, pickle
        import numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        import parity_util as pu
        from mapcaller_b200 import api, shard
        rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
        torch.cuda.set_device(rank); dist.init_process_group('nccl')
        case = pu.make_case(**%r)
        nsb = %d
        paired = %d
        ix = pu.build_index(case)
        ctx = api.Context(ix, paired=paired, device=rank, want_alignments=1, shard_rank=rank, shard_count=world)
        uid = [api.Context.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        parts = []
        for b in range(nsb):
            sseq, soff = shard.take_shard(case['seq'], case['off'], nsb, b, bool(paired))        # super-batch b of the library ...
            seq, off = shard.take_shard(sseq, soff, world, rank, bool(paired))                   # ... and this rank's share of it
            res = ctx.map_batch(seq, off)
            parts.append(dict(reads=api.unpack_reads(res), est=[int(x) for x in res['chunks']['est_distance']] if paired else [], replays=res['replays']))
        totals = ctx.totals()
        ctx.profile_allreduce()
        ins, dele = ctx.indels()
        out = dict(parts=parts, totals=totals, profile=ctx.profile_columns(), ins=ins, dele=dele, bp=ctx.breakpoints(),
                   inv=sorted(ctx.sites(0), key=lambda x: x[0]), tnl=sorted(ctx.sites(1), key=lambda x: x[0]))
        pickle.dump(out, open(%r + '/o%%d.pkl' %% rank, 'wb'))
        dist.destroy_process_group()
    """) % (ROOT, os.path.join(ROOT, "tests"), dict(CASE, paired=paired), super_batches, paired, str(tmp_path))
    script = tmp_path / "w.py"
    script.write_text(code)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29612", str(script)])
    import pickle
    got = [pickle.load(open(str(tmp_path / ("o%d.pkl" % r)), "rb")) for r in range(2)]
    case = pu.make_case(**dict(CASE, paired=paired))
    ix = pu.build_index(case)
    want = pu.oracle_results(case, ix)
    reads, est = [], []
    for b in range(super_batches):
        for r in range(2):
            reads += got[r]["parts"][b]["reads"]; est += got[r]["parts"][b]["est"]
    for r in range(2):
        t = got[r]["totals"]
        mine = dict(reads=reads, est=est, profile=got[r]["profile"], ins=got[r]["ins"], dele=got[r]["dele"], bp=got[r]["bp"], inv=got[r]["inv"], tnl=got[r]["tnl"],
                    counters=dict(reads=t["total_reads"], mapped=t["total_mapped"], paired=t["total_paired"], dist_sum=t["total_distance"], len_sum=t["read_length_sum"],
                                  avgDist=t["avg_dist"]))
        pu.assert_same(mine, want, paired=bool(paired))


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_one_process_two_devices(built):
    """Two contexts on two GPUs inside ONE process, used alternately: every entry point selects its context's device, and the
    per-device kernel attributes (dynamic shared memory of the rescue / small-fill kernels) are set on both."""
    from mapcaller_b200 import api
    case = pu.make_case(seed=63, n_pairs=5000, genome_len=90000, sv=3.0, n_dup=10, tandem=5)
    ix = pu.build_index(case)
    want = pu.oracle_results(case, ix)
    seq, off = case["seq"], case["off"]
    n = len(off) - 1
    half = (n // 400) * 200
    ctxs = [api.Context(ix, want_alignments=1, update_profile=1, device=d, **case["params"]) for d in (0, 1)]
    got = [dict(reads=[], est=[]) for _ in ctxs]
    for b, e in ((0, half), (half, n)):
        for k, ctx in enumerate(ctxs):                       # interleaved: device 0, device 1, device 0, device 1
            res = ctx.map_batch(seq[off[b]:off[e]], off[b:e + 1] - off[b])
            got[k]["reads"] += api.unpack_reads(res); got[k]["est"] += [int(x) for x in res["chunks"]["est_distance"]]
    for k, ctx in enumerate(ctxs):
        t = ctx.totals()
        got[k]["counters"] = dict(reads=t["total_reads"], mapped=t["total_mapped"], paired=t["total_paired"], dist_sum=t["total_distance"],
                                  len_sum=t["read_length_sum"], avgDist=t["avg_dist"])
        got[k]["profile"] = ctx.profile_columns(); got[k]["ins"], got[k]["dele"] = ctx.indels(); got[k]["bp"] = ctx.breakpoints()
        got[k]["inv"] = sorted(ctx.sites(0), key=lambda x: x[0]); got[k]["tnl"] = sorted(ctx.sites(1), key=lambda x: x[0])
        got[k]["summary"] = ctx.profile_summary()
        pu.assert_same(got[k], want)
    for ctx in ctxs:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("n_pairs,seed", [(20000, 64), (600, 65), (250, 66)])
def test_reduce_scatter_read_out_equals_one_gpu(built, tmp_path, n_pairs, seed):
    """mc_profile_reduce_scatter: every rank keeps the counters of one genome tile; the packed columns of the tiles, the variant
    scan (default, monomorphic and low-threshold parameters; thin libraries put gap runs across the tile border), the summary
    and the checksum must equal those of ONE context that mapped the whole library."""
    code = textwrap.dedent("""
        import os, sys
        import numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        import parity_util as pu
        from mapcaller_b200 import api, shard
        rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
        torch.cuda.set_device(rank); dist.init_process_group('nccl')
        case = pu.make_case(seed=%d, n_pairs=%d, genome_len=120000, contigs=2, sv=2.0, n_dup=10)
        ix = pu.build_index(case)
        ctx = api.Context(ix, paired=1, device=rank, shard_rank=rank, shard_count=world)
        uid = [api.Context.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        seq, off = shard.take_shard(case['seq'], case['off'], world, rank)
        ctx.map_batch(seq, off)
        ctx.profile_reduce_scatter()
        beg, end = ctx.profile_owned()
        assert (beg, end) == ((0, 76800) if rank == 0 else (76800, 120000)), (beg, end)
        solo = api.Context(ix, paired=1, device=rank)
        solo.map_batch(case['seq'], case['off'])
        assert np.array_equal(ctx.profile(beg, end), solo.profile(beg, end)), 'packed columns of the tile'
        assert ctx.profile_summary() == solo.profile_summary(), 'summary'
        assert ctx.profile_checksum() == solo.profile_checksum(), 'checksum'
        for kw in (dict(), dict(monomorphic=1), dict(min_allele_depth=1, frequency_thr=0.05, min_cnv_size=5, min_unmapped_size=5), dict(somatic=1, ploidy=1)):
            (ra, da), (rb, db) = ctx.variant_scan(**kw), solo.variant_scan(**kw)
            assert len(ra) == len(rb) and ra == rb, ('records', kw, len(ra), len(rb), [x for x, y in zip(ra, rb) if x != y][:2])
            assert np.array_equal(da, db), ('block depths', kw)
            assert len(rb) > 10
        try:
            ctx.variant_scan(gvcf=1); raise SystemExit('a gVCF scan must be refused after a reduce-scatter')
        except api.McError:
            pass
        dist.barrier(); dist.destroy_process_group()
    """) % (ROOT, os.path.join(ROOT, "tests"), seed, n_pairs)
    script = tmp_path / "w.py"
    script.write_text(code)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)])
