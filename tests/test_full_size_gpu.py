"""Parity at the sizes BASELINE.json names, against the UNMODIFIED reference (oracle/_ref, one thread = its deterministic run):

* configs[1] at full size - 4.6 Mbp genome, 1.15 M 2x100 bp pairs (50x): per-read records of the first 100 k reads, EstiDistance of
  every chunk, totals, the whole profile, indel maps, break points and SV sites of the whole library;
* configs[2] genome (248,956,422 bp, repeat families; the index lives in HBM, not in L2) with a 100 k-pair prefix of the
  library: the same, the profile compared as the raw 16-byte MappingRecord_t image of all 249 M columns.

Slow by the standards of this suite (about a minute each, nearly all of it the reference on one host thread)."""
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import parity_util as pu
from mapcaller_b200 import api, simulate as sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.slow, pytest.mark.skipif(not pu.have_ref(), reason="oracle/_ref not built")]

WORKER = textwrap.dedent("""
    import sys, pickle
    import numpy as np
    sys.path.insert(0, %r)
    import ref_oracle as ro
    job = np.load(sys.argv[1], allow_pickle=True)
    ro.load(str(job['prefix'])); ro.set_params(threads=1)
    seq, off = job['seq'], job['off']; k = int(job['parse_reads']); n = len(off) - 1
    reads, est = ro.map_reads(seq[:off[k]], off[:k + 1], True, True)          # parsed: the per-read comparison
    if n > k:                                                                  # the rest of the library: only what it leaves behind
        r2, e2 = ro.map_reads(seq[off[k]:], off[k:] - off[k], True, True); est = est + e2; del r2
    ro.lib().mcref_finish_sites()
    out = dict(reads=reads, est=est, counters=ro.counters(), ins=ro.indels(0), dele=ro.indels(1), bp=ro.breakpoints(), inv=ro.sites(0), tnl=ro.sites(1))
    if int(job['raw_profile']):
        G = ro.lib().mcref_genome_size()
        for b in range(0, G, 1 << 24):
            np.save(str(job['prefix']) + '.prof%%d.npy' %% (b >> 24), ro.profile(b, min(G, b + (1 << 24))).astype(np.uint16))
    else:
        out['profile'] = ro.profile()
    pickle.dump(out, open(sys.argv[2], 'wb'), protocol=4)
""") % os.path.join(ROOT, "tests")


def _reference(tmp_path, ix, seq, off, parse_reads, raw_profile):
    prefix = str(tmp_path / "idx"); ix.save(prefix)
    job, outp, script = str(tmp_path / "job.npz"), str(tmp_path / "out.pkl"), str(tmp_path / "w.py")
    np.savez(job, prefix=prefix, seq=seq, off=off, parse_reads=parse_reads, raw_profile=int(raw_profile))
    open(script, "w").write(WORKER)
    subprocess.run([sys.executable, script, job, outp], check=True)
    return pickle.load(open(outp, "rb")), prefix


def _mine(ix, seq, off, batch_reads, parse_reads):
    ctx = api.Context(ix, paired=1, want_alignments=1, update_profile=1)
    n = len(off) - 1
    reads, est = [], []
    cuts = [0] + list(range(min(parse_reads, n), n, batch_reads)) + [n]      # the first batch is exactly the part that is compared read by read
    for b, e in zip(cuts[:-1], cuts[1:]):
        if e <= b:
            continue
        res = ctx.map_batch(seq[off[b]:off[e]], off[b:e + 1] - off[b])
        if b < parse_reads:
            reads += api.unpack_reads(res)
        est += [int(x) for x in res["chunks"]["est_distance"]]
    t = ctx.totals()
    ins, dele = ctx.indels()
    out = dict(reads=reads[:parse_reads], est=est, ins=ins, dele=dele, bp=ctx.breakpoints(), inv=sorted(ctx.sites(0), key=lambda x: x[0]), tnl=sorted(ctx.sites(1), key=lambda x: x[0]),
               counters=dict(reads=t["total_reads"], mapped=t["total_mapped"], paired=t["total_paired"], dist_sum=t["total_distance"], len_sum=t["read_length_sum"], avgDist=t["avg_dist"]))
    return ctx, out


def _same_but_profile(mine, ref):
    d = pu.first_read_diff(mine["reads"], ref["reads"], paired=True)
    assert d is None, "read %d differs:\n mine %r\n ref  %r" % d
    assert mine["est"] == ref["est"], "EstiDistance trajectory differs"
    for k in ("reads", "mapped", "paired", "dist_sum", "len_sum", "avgDist"):
        assert mine["counters"][k] == ref["counters"][k], (k, mine["counters"][k], ref["counters"][k])
    assert mine["ins"] == ref["ins"] and mine["dele"] == ref["dele"], "indel maps differ"
    assert mine["bp"] == ref["bp"], "BreakPointMap differs"
    assert sorted(mine["inv"]) == sorted(ref["inv"]) and sorted(mine["tnl"]) == sorted(ref["tnl"]), "SV sites differ"


def test_configs1_full_size_equals_the_reference(built, tmp_path):
    case = pu.make_case(seed=7, n_pairs=1_150_000, read_len=100, genome_len=4_600_000, n_dup=200, tandem=20, sv=1.0 / 30)
    ix = pu.build_index(case)
    K = 100_000
    ctx, mine = _mine(ix, case["seq"], case["off"], 400_000, K)
    ref, _ = _reference(tmp_path, ix, case["seq"], case["off"], K, False)
    _same_but_profile(mine, ref)
    a, o = ctx.profile_columns(), ref["profile"]
    bad = np.nonzero((a != o).any(axis=1))[0]
    assert len(bad) == 0, "profile differs at column %d: %r vs %r" % (bad[0], a[bad[0]], o[bad[0]])
    assert int((o[:, :4].sum(axis=1) > 0).sum()) > 4_000_000
    ctx.close()


def test_configs2_genome_prefix_equals_the_reference(built, tmp_path):
    G, P = 248_956_422, 100_000
    g = sim.genome(G, 13, n_dup=2000, repeat_frac=0.15)
    mut, _ = sim.mutate(g, 14, snp_per_mb=1000, small_indel_per_mb=100, large_indel_per_mb=0, sv_per_mb=0)
    r1, r2 = sim.simulate_pairs_fast(mut, 250_000, 150, seed=15, frag_mean=450, frag_sd=50, sub_rate=0.003)
    seq, off = sim.interleave(r1[:P], r2[:P])
    del mut
    ix = api.Index.build(sim.encode(g), gpu_device=0)
    ctx, mine = _mine(ix, seq, off, 80_000, 2 * P)
    assert ctx.stats()["seed_blocks"] > 0
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    import tempfile, pathlib
    with tempfile.TemporaryDirectory(dir=shm) as td:
        ref, prefix = _reference(pathlib.Path(td), ix, seq, off, 2 * P, True)
        _same_but_profile(mine, ref)
        covered = 0
        for b in range(0, G, 1 << 24):
            e = min(G, b + (1 << 24))
            a, o = ctx.profile_columns(b, e), np.load(prefix + ".prof%d.npy" % (b >> 24))
            bad = np.nonzero((a != o).any(axis=1))[0]
            assert len(bad) == 0, "profile differs at column %d: %r vs %r" % (b + bad[0], a[bad[0]], o[bad[0]])
            covered += int((o[:, :4].sum(axis=1) > 0).sum())
        assert covered > 20_000_000
    ctx.close()
