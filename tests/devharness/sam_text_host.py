"""Developer harness: mc_sam_text (SAM lines assembled by the stage bodies, here compiled for the host) against the SAM of the
reference CLI.  Not a test, not a fallback."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mapcaller_b200 import api
api._LIB_PATH = os.path.join(ROOT, "tools", "hostemu", "_build", "libmc_hostemu.so")
import parity_util as pu

CASES = {
    "pe": dict(seed=3, n_pairs=3000, genome_len=60000),
    "pe_multi": dict(seed=5, n_pairs=6000, genome_len=200000, contigs=3, repeat_frac=0.3),
    "pe_ksw2": dict(seed=6, n_pairs=3000, genome_len=100000, alg_ksw2=1, indel_rate=0.002),
    "se": dict(seed=7, n_pairs=3000, genome_len=80000, paired=0),
    "lower_n": dict(seed=15, n_pairs=3000, genome_len=60000, lower_rate=0.3, n_rate=0.01),
    "all_best": dict(seed=9, n_pairs=4000, genome_len=60000, n_dup=40, tandem=10, all_best=True),
    "batched": dict(seed=12, n_pairs=5000, genome_len=100000, batch_pairs=1200),
}
for name in (sys.argv[1:] or list(CASES)):
    kw = dict(CASES[name]); all_best = kw.pop("all_best", False); bp = kw.pop("batch_pairs", None)
    case = pu.make_case(**kw)
    paired = bool(case["params"]["paired"])
    with tempfile.TemporaryDirectory() as td:
        ref = pu.sam_comparable(pu.sam_lines_reference(case, td, all_best=all_best), paired)
    mine = pu.sam_comparable(pu.sam_text_cuda(case, batch_pairs=bp, all_best=all_best), paired)
    bad = [i for i, (a, b) in enumerate(zip(mine, ref)) if a != b]
    print(name, "OK" if mine == ref else "FAIL", len(mine), len(ref), "first diff", bad[:1])
    if bad:
        print("  mine", mine[bad[0]][:300]); print("  ref ", ref[bad[0]][:300])
