"""Parity at BASELINE.json configs[2] scale: chr1-sized (248 Mbp) synthetic genome with repeat families, 2x150 bp pairs.
The CUDA path (index in HBM, compact layout, several batches) against the CPU restatement of the reference (oracle/restate),
record for record: per-read candidates / fragments / alignment strings, EstiDistance per chunk, totals, the whole profile
(tile by tile), indel maps, break points, SV sites.  Test infrastructure like the rest of tests/ (it is the only kind of code that may use oracle/), but not collected by pytest (takes a few minutes and ~20 GB of
host memory); its log is committed under profiles/.
usage: python tests/scale_parity.py [genome_bp] [pairs] [read_len] [batches] [ksw2]
(`ksw2` as fifth argument = configs[4] flavour: -alg ksw2, small indels 2000/Mb, large 500/Mb, 0.2 % indel errors per read base)"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mapcaller_b200 import api, simulate as sim
import cpu_oracle
import parity_util as pu

G = int(sys.argv[1]) if len(sys.argv) > 1 else 248_956_422
P = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
L = int(sys.argv[3]) if len(sys.argv) > 3 else 150
NB = int(sys.argv[4]) if len(sys.argv) > 4 else 3
KSW2 = len(sys.argv) > 5 and sys.argv[5] == "ksw2"
t = time.time()
g = sim.genome(G, 13, n_dup=2000, repeat_frac=0.15 if G > 50_000_000 else 0.0)
if KSW2:
    mut, _ = sim.mutate(g, 14, snp_per_mb=1000, small_indel_per_mb=2000, large_indel_per_mb=500, sv_per_mb=1)
    r1, r2 = sim.simulate_pairs(mut[3000:], P, L, seed=15, frag_mean=max(450, 2 * L), frag_sd=50, sub_rate=0.003, indel_rate=0.002)
else:
    mut, _ = sim.mutate(g, 14, snp_per_mb=1000, small_indel_per_mb=100, large_indel_per_mb=20, sv_per_mb=1)
    r1, r2 = sim.simulate_pairs(mut[3000:], P, L, seed=15, frag_mean=450, frag_sd=50, sub_rate=0.003)
seq, off = sim.interleave(r1, r2)
del mut
print("data %.1fs" % (time.time() - t), flush=True)
# contigs of at most 2^31 - 1 bases (chromosome lengths are int32 in the reference's .ann); above 2^32 text symbols the
# context keeps the reference's 128-row index blocks and 64-bit rows (the path of a GRCh38-sized genome)
NC = max(1, (G + 249_999_999) // 250_000_000)
lens = [G // NC] * NC; lens[-1] += G - sum(lens)
t = time.time(); ix = api.Index.build(sim.encode(g), chrom_len=lens, chrom_name=["ctg%d" % (i + 1) for i in range(NC)])
print("index build %.1fs (%d contigs)" % (time.time() - t, NC), flush=True)
params = dict(paired=1, alg_ksw2=int(KSW2))
n = len(off) - 1
mine = dict(reads=[], est=[], replays=0)
t = time.time()
with api.Context(ix, want_alignments=1, update_profile=1, **params) as ctx:
    per = ((n // 200 + NB - 1) // NB) * 200
    for b in range(0, n, per):
        e = min(n, b + per)
        res = ctx.map_batch(seq[off[b]:off[e]], off[b:e + 1] - off[b])
        mine["reads"] += api.unpack_reads(res); mine["est"] += [int(x) for x in res["chunks"]["est_distance"]]; mine["replays"] += res["replays"]
    tt = ctx.totals()
    mine["counters"] = dict(reads=tt["total_reads"], mapped=tt["total_mapped"], paired=tt["total_paired"], dist_sum=tt["total_distance"],
                            len_sum=tt["read_length_sum"], avgDist=tt["avg_dist"])
    print("cuda %.1fs, replays %d, totals %r" % (time.time() - t, mine["replays"], mine["counters"]), flush=True)
    with tempfile.TemporaryDirectory() as td:
        prefix = os.path.join(td, "idx"); ix.save(prefix)
        t = time.time()
        orc = cpu_oracle.Oracle(prefix, **params)
        reads, est = orc.map_reads(seq, off, True, True)
        print("oracle %.1fs" % (time.time() - t), flush=True)
        d = pu.first_read_diff(mine["reads"], reads, paired=True)
        assert d is None, "read %d differs:\n mine %r\n ref  %r" % d
        assert mine["est"] == est, "EstiDistance trajectory differs"
        oc = orc.counters()
        for k in ("reads", "mapped", "paired", "dist_sum", "len_sum", "avgDist"):
            assert mine["counters"][k] == oc[k], (k, mine["counters"][k], oc[k])
        tile = 1 << 24; ncol = 0
        for b in range(0, G, tile):
            e = min(G, b + tile)
            a, o = ctx.profile_columns(b, e), orc.profile(b, e)
            bad = np.nonzero((a != o).any(axis=1))[0]
            assert len(bad) == 0, "profile differs at column %d: %r vs %r" % (b + bad[0], a[bad[0]], o[bad[0]])
            ncol += int((o[:, :4].sum(axis=1) > 0).sum())
        ins, dele = ctx.indels()
        assert ins == orc.indels(0) and dele == orc.indels(1), "indel maps differ"
        assert ctx.breakpoints() == orc.breakpoints(), "BreakPointMap differs"
        assert sorted(ctx.sites(0)) == sorted(orc.sites(0)) and sorted(ctx.sites(1)) == sorted(orc.sites(1)), "SV sites differ"
        w = orc.work(); st = ctx.stats()
        assert 0.985 * w["seed_blocks"] <= st["seed_blocks"] <= w["seed_blocks"] and st["sa_reads"] == w["sa_reads"], "work counters differ"
        print("PARITY OK: %d reads, %d chunks, %d covered columns, %d insertions, %d deletions, %d break points, %d inversion sites, %d translocation sites"
              % (n, len(est), ncol, len(ins), len(dele), len(ctx.breakpoints()), len(ctx.sites(0)), len(ctx.sites(1))), flush=True)
        orc.close()
