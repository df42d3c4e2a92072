"""Randomised cases through TWO GPUs on one library (ordered exchange) against the unmodified reference's single run.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29677 tests/devharness/fuzz_two_gpus.py [seed] [cases]
Developer harness (uses the oracle); not collected by pytest."""
import os, sys, pickle, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
import parity_util as pu
from mapcaller_b200 import api, shard

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dist.init_process_group("nccl")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
tmp = [tempfile.mkdtemp() if rank == 0 else None]; dist.broadcast_object_list(tmp, src=0); tmp = tmp[0]
bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    rl = int(rng.choice([50, 100, 150, 250]))
    kw = dict(seed=int(rng.integers(100, 10000)), n_pairs=int(rng.integers(3000, 9000)), genome_len=int(rng.choice([30000, 120000])), read_len=rl,
              contigs=int(rng.choice([1, 3])), sub_rate=float(rng.choice([0.002, 0.02])), indel_rate=float(rng.choice([0, 0.004])), n_rate=float(rng.choice([0, 0.004])),
              sv=float(rng.choice([0, 4])), n_dup=int(rng.choice([0, 12])), tandem=int(rng.choice([0, 6])), frag_mean=float(max(rl + 30, rng.choice([250, 420]))), frag_sd=float(rng.choice([15, 70])),
              paired=int(rng.choice([1, 1, 0])), alg_ksw2=int(rng.choice([0, 1])), max_dup=int(rng.choice([2, 5, 15])))
    nsb = int(rng.choice([1, 2, 4]))
    paired = kw["paired"]
    case = pu.make_case(**kw); ix = pu.build_index(case)
    ctx = api.Context(ix, paired=paired, alg_ksw2=kw["alg_ksw2"], max_dup=kw["max_dup"], device=rank, want_alignments=1, shard_rank=rank, shard_count=world)
    uid = [api.Context.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    parts = []
    for b in range(nsb):
        sseq, soff = shard.take_shard(case["seq"], case["off"], nsb, b, bool(paired))
        seq, off = shard.take_shard(sseq, soff, world, rank, bool(paired))
        res = ctx.map_batch(seq, off)
        parts.append(dict(reads=api.unpack_reads(res), est=[int(x) for x in res["chunks"]["est_distance"]] if paired else []))
    totals = ctx.totals(); ctx.profile_allreduce(); ins, dele = ctx.indels()
    out = dict(parts=parts, totals=totals, profile=ctx.profile_columns(), ins=ins, dele=dele, bp=ctx.breakpoints(), inv=sorted(ctx.sites(0)), tnl=sorted(ctx.sites(1)))
    pickle.dump(out, open(os.path.join(tmp, "o%d_%d.pkl" % (it, rank)), "wb"))
    ctx.close(); dist.barrier()
    if rank == 0:
        got = [pickle.load(open(os.path.join(tmp, "o%d_%d.pkl" % (it, r)), "rb")) for r in range(world)]
        reads, est = [], []
        for b in range(nsb):
            for r in range(world):
                reads += got[r]["parts"][b]["reads"]; est += got[r]["parts"][b]["est"]
        want = pu.ref_results(case, ix) if pu.have_ref() else pu.oracle_results(case, ix)
        try:
            for r in range(world):
                t = got[r]["totals"]
                mine = dict(reads=reads, est=est, profile=got[r]["profile"], ins=got[r]["ins"], dele=got[r]["dele"], bp=got[r]["bp"], inv=got[r]["inv"], tnl=got[r]["tnl"],
                            counters=dict(reads=t["total_reads"], mapped=t["total_mapped"], paired=t["total_paired"], dist_sum=t["total_distance"], len_sum=t["read_length_sum"], avgDist=t["avg_dist"]))
                pu.assert_same(mine, want, paired=bool(paired))
            print("OK", it, "super-batches", nsb, kw, flush=True)
        except AssertionError as e:
            bad += 1; print("FAIL", it, nsb, kw, repr(e)[:300], flush=True)
    dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
