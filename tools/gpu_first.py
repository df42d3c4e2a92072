"""First GPU bring-up: parity of the CUDA path against oracle/_ref on a few seeded cases + rough timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_util as pu
from mapcaller_b200 import api, simulate as sim

for name, kw in [("small", dict(seed=3, n_pairs=600, genome_len=40000)), ("mid", dict(seed=5, n_pairs=20000, genome_len=300000, contigs=3)),
                 ("ksw2", dict(seed=6, n_pairs=5000, genome_len=100000, alg_ksw2=1, indel_rate=0.002)), ("se", dict(seed=7, n_pairs=4000, genome_len=80000, paired=0)),
                 ("nbase", dict(seed=8, n_pairs=5000, genome_len=80000, n_rate=0.01, sub_rate=0.02)), ("sv", dict(seed=9, n_pairs=20000, genome_len=200000, sv=5.0, n_dup=30, tandem=20)),
                 ("long", dict(seed=10, n_pairs=4000, genome_len=150000, read_len=250, frag_mean=600, frag_sd=80, indel_rate=0.003)), ("dup", dict(seed=11, n_pairs=30000, genome_len=20000, max_dup=3))]:
    case = pu.make_case(**kw); ix = pu.build_index(case)
    t = time.time(); mine = pu.cuda_results(case, ix); t1 = time.time() - t
    ref = pu.ref_results(case, ix)
    try:
        pu.assert_same(mine, ref, paired=bool(case['params']['paired'])); print(name, "PARITY OK", "%.3fs" % t1, "replays", mine["replays"], mine["stats"], flush=True)
    except AssertionError as e:
        print(name, "PARITY FAIL", str(e)[:2000], flush=True)

# rough throughput at E. coli size
g = sim.genome(4_600_000, 7, n_dup=200)
mut, _ = sim.mutate(g, 8)
r1, r2 = sim.simulate_pairs(mut, 1_150_000, 100, seed=11)
seq, off = sim.interleave(r1, r2)
t = time.time(); ix = api.Index.build(sim.encode(g)); print("index build %.2fs" % (time.time() - t), flush=True)
with api.Context(ix, paired=1) as ctx:
    for it in range(3):
        ctx.reset_stats(); t = time.time(); res = ctx.map_batch(seq, off); dt = time.time() - t
        print("iter", it, "%.3fs" % dt, "pairs/s %.0f" % (len(r1) / dt), "replays", res["replays"], ctx.stats(), flush=True)
    print(ctx.totals())
