"""Developer harness: parity cases through the host-compiled stage bodies (tools/hostemu/build.sh) against oracle/_ref on a
box without a GPU.  Lives under tests/ because it uses the oracle; not collected by pytest, never imported elsewhere."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mapcaller_b200 import api
api._LIB_PATH = os.path.join(ROOT, "tools", "hostemu", "_build", "libmc_hostemu.so")
import parity_util as pu

CASES = {
    "small": dict(seed=3, n_pairs=600, genome_len=40000),
    "mid": dict(seed=5, n_pairs=20000, genome_len=300000, contigs=3),
    "ksw2": dict(seed=6, n_pairs=5000, genome_len=100000, alg_ksw2=1, indel_rate=0.002),
    "se": dict(seed=7, n_pairs=4000, genome_len=80000, paired=0),
    "nbase": dict(seed=8, n_pairs=5000, genome_len=80000, n_rate=0.01, sub_rate=0.02),
    "sv": dict(seed=9, n_pairs=20000, genome_len=200000, sv=5.0, n_dup=30, tandem=20),
    "long": dict(seed=10, n_pairs=4000, genome_len=150000, read_len=250, frag_mean=600, frag_sd=80, indel_rate=0.003),
    "dup": dict(seed=11, n_pairs=30000, genome_len=20000, max_dup=3),
    "lower": dict(seed=15, n_pairs=4000, genome_len=60000, lower_rate=0.2),
    "batched": dict(seed=12, n_pairs=12000, genome_len=100000),
}
names = sys.argv[1:] or list(CASES)
for name in names:
    kw = CASES[name]
    case = pu.make_case(**kw); ix = pu.build_index(case)
    t = time.time(); mine = pu.cuda_results(case, ix, batch_reads=4000 if name == "batched" else None); t1 = time.time() - t
    t = time.time(); ref = pu.ref_results(case, ix); t2 = time.time() - t
    try:
        pu.assert_same(mine, ref, paired=bool(case['params']['paired']))
        print(name, "PARITY OK  mine %.2fs ref %.2fs replays %d  dp_tasks %d ins %d del %d bp %d inv %d tnl %d" % (t1, t2, mine["replays"], mine["stats"]["dp_tasks"], len(ref["ins"]), len(ref["dele"]), len(ref["bp"]), len(ref["inv"]), len(ref["tnl"])), flush=True)
    except AssertionError as e:
        print(name, "PARITY FAIL", str(e)[:3000], flush=True)
