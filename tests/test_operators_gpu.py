"""Operator-level entries for the steps of the path that otherwise only show up inside a mapped batch:
mc_read_alignment_batch (ProduceReadAlignment in the single-end branch of ReadMapping()) and mc_rescue_batch (AlignmentRescue with a
given EstiDistance), each against the unmodified reference's own functions driven the same way (oracle/_ref), and each leaving
the context's sequential state alone."""
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import parity_util as pu
from mapcaller_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pu.have_ref(), reason="oracle/_ref not built")]

WORKER = textwrap.dedent("""
    import sys, pickle
    import numpy as np
    sys.path.insert(0, %r)
    import ref_oracle as ro
    job = np.load(sys.argv[1], allow_pickle=True)
    ro.load(str(job['prefix'])); ro.set_params(threads=1, nw=not int(job['ksw2']))
    seq, off = job['seq'], job['off']; n = len(off) - 1
    reads = []
    if int(job['avg_dist']) < 0:                      # ProduceReadAlignment: the single-end branch, read by read
        reads, _ = ro.map_reads(seq, off, False, False)
    else:                                             # AlignmentRescue: every 200-read chunk starts from the same avgDist
        for b in range(0, n, 200):
            e = min(n, b + 200)
            ro.lib().mcref_reset_state(); ro.set_avg_dist(int(job['avg_dist']))
            r, est = ro.map_reads(seq[off[b]:off[e]], off[b:e + 1] - off[b], True, False)
            assert est == [int(int(job['avg_dist']) * 1.5)], est
            reads += r
    pickle.dump(reads, open(sys.argv[2], 'wb'), protocol=4)
""") % os.path.join(ROOT, "tests")


def _reference(tmp_path, ix, seq, off, avg_dist, ksw2):
    prefix = str(tmp_path / "idx"); ix.save(prefix)
    job, outp, script = str(tmp_path / "job.npz"), str(tmp_path / "out.pkl"), str(tmp_path / "w.py")
    np.savez(job, prefix=prefix, seq=seq, off=off, avg_dist=avg_dist, ksw2=int(ksw2))
    open(script, "w").write(WORKER)
    subprocess.run([sys.executable, script, job, outp], check=True)
    return pickle.load(open(outp, "rb"))


@pytest.mark.parametrize("ksw2", [0, 1])
def test_read_alignment_operator_matches_the_reference(built, tmp_path, ksw2):
    case = pu.make_case(seed=71, n_pairs=1500, genome_len=90000, contigs=2, sv=3.0, n_dup=10, tandem=5, indel_rate=0.003, n_rate=0.002, alg_ksw2=ksw2)
    ix = pu.build_index(case)
    ctx = api.Context(ix, paired=1, alg_ksw2=ksw2, update_profile=1)            # a PAIRED context with a profile: the entry must not care
    ctx.map_batch(case["seq"], case["off"])
    before = (ctx.totals(), ctx.profile_checksum())
    n = 1001                                                                   # an odd number of reads, not a multiple of 200
    seq, off = case["seq"][:case["off"][n]], case["off"][:n + 1]
    res = ctx.read_alignment_batch(seq, off)
    ref = _reference(tmp_path, ix, seq, off, -1, ksw2)
    d = pu.first_read_diff(api.unpack_reads(res), ref, paired=False)
    assert d is None, "read %d differs:\n mine %r\n ref  %r" % d
    assert (ctx.totals(), ctx.profile_checksum()) == before                    # totals, avgDist and the profile are where they were
    ctx.map_batch(case["seq"], case["off"])                                    # ... and the library goes on
    ctx.close()


@pytest.mark.parametrize("avg_dist", [1000, 300, 180])
def test_rescue_operator_matches_the_reference(built, tmp_path, avg_dist):
    # repeats and structural variants put mates where only the rescue finds them; fragments of ~380 +- 60: with avg_dist 180
    # (EstiDistance 270) most mates fall outside the window, with 1000 all of them inside
    case = pu.make_case(seed=72, n_pairs=2000, genome_len=90000, contigs=2, sv=4.0, n_dup=20, tandem=8, frag_mean=380, frag_sd=60, sub_rate=0.02, indel_rate=0.004)
    ix = pu.build_index(case)
    ctx = api.Context(ix, paired=1, update_profile=1)
    ctx.map_batch(case["seq"], case["off"])
    before = (ctx.totals(), ctx.profile_checksum(), sorted(ctx.sites(0)), sorted(ctx.sites(1)))
    res = ctx.rescue_batch(case["seq"], case["off"], avg_dist)
    assert [int(x) for x in res["chunks"]["est_distance"]] == [int(avg_dist * 1.5)] * (len(case["off"]) // 200)
    ref = _reference(tmp_path, ix, case["seq"], case["off"], avg_dist, 0)
    mine = api.unpack_reads(res)
    d = pu.first_read_diff(mine, ref, paired=True)
    assert d is None, "read %d differs:\n mine %r\n ref  %r" % d
    assert (ctx.totals(), ctx.profile_checksum(), sorted(ctx.sites(0)), sorted(ctx.sites(1))) == before
    ctx.close()


@pytest.mark.parametrize("paired", [1, 0])
def test_deferred_profile_update_equals_the_inline_one_and_the_reference(built, paired):
    """mc_defer_profile + mc_update_profile_last: UpdateProfile / UpdateMultiHitCount applied as a step of its own to the reads of the
    batch just mapped - batch after batch the same profile, indel maps and break points as the inline update, which the
    reference's own UpdateProfile leaves behind (pu.ref_results)."""
    case = pu.make_case(seed=73, n_pairs=4000, genome_len=90000, contigs=2, sv=3.0, n_dup=15, tandem=6, indel_rate=0.003, lower_rate=0.05, paired=paired, max_dup=3)
    ix = pu.build_index(case)
    seq, off = case["seq"], case["off"]
    n = len(off) - 1
    cut = (n // 600) * 200
    inline, later = api.Context(ix, update_profile=1, **case["params"]), api.Context(ix, update_profile=1, **case["params"])
    later.defer_profile(True)
    empty = later.profile_checksum()
    for b, e in ((0, cut), (cut, n)):
        inline.map_batch(seq[off[b]:off[e]], off[b:e + 1] - off[b])
        before = later.profile_checksum()
        later.map_batch(seq[off[b]:off[e]], off[b:e + 1] - off[b])
        assert later.profile_checksum() == before                    # mapping alone leaves the profile alone
        later.update_profile_last()
        assert later.profile_checksum() == inline.profile_checksum()
        with pytest.raises(api.McError):
            later.update_profile_last()                              # once per batch
    assert empty != inline.profile_checksum()
    assert later.indels() == inline.indels() and later.breakpoints() == inline.breakpoints() and later.profile_summary() == inline.profile_summary()
    ref = pu.ref_results(case, ix, want_reads=False)
    assert np.array_equal(later.profile_columns(), ref["profile"])
    ins, dele = later.indels()
    assert ins == ref["ins"] and dele == ref["dele"] and later.breakpoints() == ref["bp"]
    inline.close(); later.close()
