"""HBM-regime experiment: a chr1-sized (default 248 Mbp) synthetic genome whose index (bwt G bytes + SA G/2 bytes) does not
fit the 126 MB L2, so the seed / locate kernels gather from HBM.  Prints the stage timers and GB/s of one resident pass.
usage: python tools/big_genome.py [genome_bp] [pairs] [read_len] [ksw2]
(`ksw2` = configs[4] flavour: -alg ksw2, indel-rich mutant and reads: the gapped-fill stress)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mapcaller_b200 import api, simulate as sim

G = int(sys.argv[1]) if len(sys.argv) > 1 else 248_956_422
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
L = int(sys.argv[3]) if len(sys.argv) > 3 else 150
KSW2 = len(sys.argv) > 4 and sys.argv[4] == "ksw2"
t = time.time(); g = sim.genome(G, 13, n_dup=2000, repeat_frac=0.15 if G > 50_000_000 else 0.0); print("genome %.1fs" % (time.time() - t), flush=True)
t = time.time()
mut, _ = sim.mutate(g, 14, snp_per_mb=1000, small_indel_per_mb=2000 if KSW2 else 100, large_indel_per_mb=500 if KSW2 else 0, sv_per_mb=0); print("mutant %.1fs" % (time.time() - t), flush=True)
t = time.time(); r1, r2 = sim.simulate_pairs(mut, P, L, seed=15, frag_mean=max(450, 2 * L), frag_sd=50, sub_rate=0.003, indel_rate=0.002 if KSW2 else 0.0)
seq, off = sim.interleave(r1, r2); print("reads %.1fs" % (time.time() - t), flush=True)
del mut
if os.environ.get("MC_GPU_INDEX"):
    t = time.time(); ix = api.Index.build(sim.encode(g), gpu_device=0); print("index build %.1fs (suffix array on the GPU)" % (time.time() - t), flush=True)
else:
    t = time.time(); ix = api.Index.build(sim.encode(g)); print("index build %.1fs (%d threads)" % (time.time() - t, os.cpu_count()), flush=True)
ctx = api.Context(ix, paired=1, alg_ksw2=int(KSW2))
ctx.stage_batch(seq, off, 0)
for i in range(3):
    ctx.reset(); ctx.reset_stats(); t = time.time(); ctx.map_staged(0); dt = time.time() - t
    st = ctx.stats()
    print(json.dumps(dict(iter=i, wall_ms=1000 * dt, pairs_per_s=P / dt, seed_ms=st["ms_seed"], seed_GBs=st["seed_blocks"] * 64 / st["ms_seed"] / 1e6,
                          locate_ms=st["ms_locate"], locate_GBs=(st["locate_blocks"] * 64 + st["sa_reads"] * 8) / st["ms_locate"] / 1e6,
                          blocks_per_pair=st["seed_blocks"] / P, dp_cells=st["dp_cells"], dp_tasks=st["dp_tasks"], stages={k: round(v, 3) for k, v in st.items() if k.startswith("ms_")}, totals=ctx.totals())), flush=True)
