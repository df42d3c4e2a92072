"""Developer harness: randomised parity of mc_variant_scan (host-compiled stage bodies, tools/hostemu/build.sh) against
CalBlockReadDepth + IdentifyVariants of the unmodified reference (oracle/_ref) - random genomes, coverages from <1x to
~80x, error / indel rates, duplicate gates and variant-calling thresholds.  Not collected by pytest.
usage: python tests/devharness/fuzz_variant_scan.py [n_cases] [first_seed]      (MC_FUZZ_GPU=1: the CUDA library instead)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mapcaller_b200 import api
if not os.environ.get("MC_FUZZ_GPU"):
    api._LIB_PATH = os.path.join(ROOT, "tools", "hostemu", "_build", "libmc_hostemu.so")
import parity_util as pu

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
bad = 0
for seed in range(seed0, seed0 + n_cases):
    rng = np.random.default_rng(seed)
    glen = int(rng.integers(15000, 150000))
    cov = float(rng.choice([0.3, 2, 8, 20, 40, 80]))
    kw = dict(seed=seed, genome_len=glen, n_pairs=max(50, int(cov * glen / 200)), contigs=int(rng.integers(1, 4)),
              sub_rate=float(rng.choice([0.0, 0.005, 0.03])), indel_rate=float(rng.choice([0.0, 0.002, 0.01])),
              n_dup=int(rng.integers(0, 20)), tandem=int(rng.integers(0, 10)), sv=float(rng.choice([0.0, 1.0, 5.0])),
              max_dup=int(rng.choice([1, 3, 5, 15])), alg_ksw2=int(rng.integers(0, 2)), paired=int(rng.random() < 0.8))
    sets = []
    for _ in range(4):
        mode = int(rng.integers(0, 3))
        sets.append(dict(min_allele_depth=int(rng.integers(1, 9)), frequency_thr=float(rng.choice([0.05, 0.2, 0.35, 0.5])),
                         somatic=int(rng.random() < 0.25), gvcf=int(mode == 1), monomorphic=int(mode == 2), ploidy=int(rng.integers(1, 3)),
                         min_cnv_size=int(rng.integers(0, 60)), min_unmapped_size=int(rng.integers(1, 60))))
    t = time.time()
    case = pu.make_case(**kw); ix = pu.build_index(case)
    mine = pu.cuda_results(case, ix, want_reads=False, vc=sets)
    ref = pu.ref_results(case, ix, want_reads=False, vc=sets)
    try:
        pu.assert_same_variants(mine, ref)
        print("seed %d OK   G=%d cov=%g  records %s  %.1fs" % (seed, glen, cov, [len(v) for v, _ in ref["vc"]], time.time() - t), flush=True)
    except AssertionError as e:
        bad += 1
        print("seed %d FAIL %s\n   case %r\n   sets %r" % (seed, str(e)[:300], kw, sets), flush=True)
print("%d of %d cases differ" % (bad, n_cases))
