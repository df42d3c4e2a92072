"""Randomised parity cases (read length, contigs, error / indel / N rates, repeats, SE / PE, nw / ksw2, gate and clip
parameters, batch cuts) through the host-compiled stage bodies against the CPU restatement and the unmodified reference.
usage: python tests/devharness/fuzz_parity.py [seed] [cases]   (developer harness; not collected by pytest)"""
import sys, traceback, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from mapcaller_b200 import api
if not os.environ.get('MC_FUZZ_GPU'):   # default: the host harness; MC_FUZZ_GPU=1 runs the CUDA library instead
    api._LIB_PATH = os.path.join(ROOT, 'tools', 'hostemu', '_build', 'libmc_hostemu.so')
import parity_util as pu
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 8):
    rl=int(rng.choice([36,50,76,100,150,250,400,700]))
    kw=dict(seed=int(rng.integers(100,10000)), n_pairs=int(rng.integers(300,2500)), genome_len=int(rng.choice([5000,20000,60000,150000])), read_len=rl,
            contigs=int(rng.choice([1,2,5,12])), sub_rate=float(rng.choice([0.001,0.01,0.03])), indel_rate=float(rng.choice([0,0.002,0.01])), n_rate=float(rng.choice([0,0.005])),
            sv=float(rng.choice([0,2,6])), n_dup=int(rng.choice([0,6,30])), tandem=int(rng.choice([0,3,20])), frag_mean=float(max(rl+20, rng.choice([200,400,700]))), frag_sd=float(rng.choice([10,40,120])),
            lower_rate=float(rng.choice([0,0.1])), paired=int(rng.choice([1,1,1,0])), alg_ksw2=int(rng.choice([0,1])), max_dup=int(rng.choice([1,5,15])), max_clip=int(rng.choice([2,5,20])),
            max_pos_diff=int(rng.choice([5,30,60])), max_mismatch_rate=float(rng.choice([0.02,0.05,0.15])))
    if kw["genome_len"]//kw["contigs"] < 3*rl+3500: kw["contigs"]=1
    if kw["genome_len"] < 4000+3*rl: kw["genome_len"]=20000
    t=time.time()
    try:
        case=pu.make_case(**kw); ix=pu.build_index(case)
        mine=pu.cuda_results(case, ix, batch_reads=int(rng.choice([200,1000,100000])))
        pu.assert_same(mine, pu.oracle_results(case, ix), paired=bool(kw["paired"]))
        ref_ok="-"
        if pu.have_ref():
            pu.assert_same(mine, pu.ref_results(case, ix), paired=bool(kw["paired"])); ref_ok="ref ok"
        print("OK", it, ref_ok, "%.1fs"%(time.time()-t), kw, flush=True)
    except BaseException as e:
        print("FAIL", it, kw, repr(e)[:300], flush=True)
