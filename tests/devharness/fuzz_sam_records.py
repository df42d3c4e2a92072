"""Developer harness: randomised parity of mc_sam_records (host-compiled stage bodies, tools/hostemu/build.sh) against the
SAM text of the unmodified reference CLI (oracle/_ref/MapCaller -t 1).  Not collected by pytest.
usage: python tests/devharness/fuzz_sam_records.py [n_cases] [first_seed]      (MC_FUZZ_GPU=1: the CUDA library instead)"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mapcaller_b200 import api
if not os.environ.get("MC_FUZZ_GPU"):
    api._LIB_PATH = os.path.join(ROOT, "tools", "hostemu", "_build", "libmc_hostemu.so")
import parity_util as pu

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
bad_cases = 0
for seed in range(seed0, seed0 + n_cases):
    rng = np.random.default_rng(seed)
    rl = int(rng.choice([60, 100, 150, 250]))
    kw = dict(seed=seed, genome_len=int(rng.integers(30000, 150000)), n_pairs=int(rng.integers(500, 5000)), contigs=int(rng.integers(1, 5)), read_len=rl,
              frag_mean=float(rng.choice([2.2, 3, 4])) * rl, frag_sd=0.3 * rl,
              sub_rate=float(rng.choice([0.0, 0.005, 0.03])), indel_rate=float(rng.choice([0.0, 0.002, 0.01])), n_rate=float(rng.choice([0.0, 0.005])),
              n_dup=int(rng.integers(0, 25)), tandem=int(rng.integers(0, 15)), sv=float(rng.choice([0.0, 2.0, 6.0])),
              repeat_frac=float(rng.choice([0.0, 0.0, 0.4])), repeat_div=float(rng.choice([0.005, 0.02])),
              max_dup=int(rng.choice([1, 5, 15])), max_clip=int(rng.choice([2, 5, 12])), max_mismatch_rate=float(rng.choice([0.03, 0.05, 0.1])),
              alg_ksw2=int(rng.integers(0, 2)), paired=int(rng.random() < 0.75))
    t = time.time()
    case = pu.make_case(**kw); paired = bool(kw["paired"])
    n = len(case["off"]) - 1
    batch = None if rng.random() < 0.5 else int(rng.integers(1, 6)) * 400
    import subprocess
    with tempfile.TemporaryDirectory() as td:
        try:
            ref = pu.sam_comparable(pu.sam_lines_reference(case, td), paired)
        except subprocess.CalledProcessError as e:   # the reference's own undefined behaviour: nothing to compare with
            print("seed %d SKIP the reference CLI exited with %d" % (seed, e.returncode), flush=True)
            continue
        mine = pu.sam_comparable(pu.sam_lines_cuda(case, batch_reads=batch), paired)
    bad = [k for k, (a, b) in enumerate(zip(mine, ref)) if a != b]
    if bad or len(mine) != len(ref):
        bad_cases += 1
        print("seed %d FAIL: %d of %d / %d lines differ\n   %r\n   %r\n   case %r" % (seed, len(bad), len(mine), len(ref), mine[bad[0]][:250] if bad else b"", ref[bad[0]][:250] if bad else b"", kw), flush=True)
    else:
        print("seed %d OK   %d lines, %s, rlen %d, batch %s  %.1fs" % (seed, len(ref), "PE" if paired else "SE", rl, batch, time.time() - t), flush=True)
print("%d of %d cases differ" % (bad_cases, n_cases))
