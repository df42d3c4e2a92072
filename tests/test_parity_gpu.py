"""Parity of the CUDA path (through the C ABI) with the oracle: bit-exact per-read records, EstiDistance trajectory,
totals, profile, indel maps, break points and SV site lists.

Checker used, in order of availability on the box: oracle/_ref (the unmodified reference; only for cases that
stay clear of its undefined behaviour), the CPU restatement oracle/libmcoracle.so (pinned against the
reference by tests/test_oracle.py), and the committed golden vectors."""
import os
import random
import tempfile

import numpy as np
import pytest

import cpu_oracle
import golden_util as gu
import parity_util as pu

pytestmark = pytest.mark.gpu

CASES = {
    "small": dict(seed=3, n_pairs=600, genome_len=40000),
    "multi_contig_sv": dict(seed=5, n_pairs=20000, genome_len=300000, contigs=3),
    "ksw2_indels": dict(seed=6, n_pairs=5000, genome_len=100000, alg_ksw2=1, indel_rate=0.002),
    "single_end": dict(seed=7, n_pairs=4000, genome_len=80000, paired=0),
    "n_bases_noisy": dict(seed=8, n_pairs=5000, genome_len=80000, n_rate=0.01, sub_rate=0.02),
    "sv_repeats": dict(seed=9, n_pairs=20000, genome_len=200000, sv=5.0, n_dup=30, tandem=20),
    "long_reads_250": dict(seed=10, n_pairs=4000, genome_len=150000, read_len=250, frag_mean=600, frag_sd=80, indel_rate=0.003),
    "long_reads_900": dict(seed=27, n_pairs=600, genome_len=150000, read_len=900, frag_mean=2200, frag_sd=200, indel_rate=0.002, sv=3.0),
    "long_reads_900_ksw2": dict(seed=28, n_pairs=400, genome_len=150000, read_len=900, frag_mean=2200, frag_sd=200, indel_rate=0.003, alg_ksw2=1),
    "deep_duplicates": dict(seed=11, n_pairs=30000, genome_len=20000, max_dup=3),
    "lower_case_reads": dict(seed=15, n_pairs=4000, genome_len=60000, lower_rate=0.2),
    "params": dict(seed=13, n_pairs=3000, genome_len=60000, max_pos_diff=8, max_clip=2, max_dup=15, max_mismatch_rate=0.1),
}


@pytest.mark.parametrize("name", sorted(gu.GOLDEN_CASES))
def test_cuda_matches_golden(built, name):
    """Committed outputs of the unmodified reference."""
    case, ref = gu.load(name)
    ix = pu.build_index(case)
    mine = pu.cuda_results(case, ix)
    pu.assert_same(mine, ref, paired=bool(case["params"]["paired"]))
    assert mine["stats"]["kernel_launches"] > 0


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_oracle(built, name):
    case = pu.make_case(**CASES[name])
    ix = pu.build_index(case)
    mine = pu.cuda_results(case, ix)
    orc = pu.oracle_results(case, ix)
    pu.assert_same(mine, orc, paired=bool(case["params"]["paired"]))
    # the work counter the seed-kernel roofline is computed from is the oracle's: exactly while a seed is searched step by
    # step or through the k-mer table, and one block per compared base once a seed is down to a single occurrence (the
    # reference reads two when the row straddles a block boundary, once in 128 steps - that share is not counted).
    # The locate kernel meets a sampled row after 3 LF steps on average (every 4th row is sampled in HBM) where
    # the reference needs 31
    assert mine["stats"]["sa_reads"] == orc["work"]["sa_reads"]
    assert 0.985 * orc["work"]["seed_blocks"] <= mine["stats"]["seed_blocks"] <= orc["work"]["seed_blocks"]
    assert mine["stats"]["locate_blocks"] < 0.25 * orc["work"]["locate_blocks"]


@pytest.mark.skipif(not pu.have_ref(), reason="oracle/_ref not on this box")
@pytest.mark.parametrize("name", ["small", "ksw2_indels", "single_end", "n_bases_noisy"])
def test_cuda_matches_reference(built, name):
    case = pu.make_case(**CASES[name])
    ix = pu.build_index(case)
    pu.assert_same(pu.cuda_results(case, ix), pu.ref_results(case, ix), paired=bool(case["params"]["paired"]))


@pytest.mark.parametrize("batch_reads", [200, 1000, 7400])
def test_batch_split_invariance(built, batch_reads):
    """The result must not depend on how the read stream is cut into batches (chunk protocol + avgDist feedback)."""
    case = pu.make_case(seed=12, n_pairs=12000, genome_len=100000)
    ix = pu.build_index(case)
    whole = pu.cuda_results(case, ix)
    parts = pu.cuda_results(case, ix, batch_reads=batch_reads)
    pu.assert_same(parts, whole)


def test_index_layouts_agree(built):
    """Texts shorter than 2^32 symbols use the compact device index (32-byte blocks of 64 rows); longer ones the reference's
    128-row blocks and a 64-bit suffix-array sample of every 8th row (mc_params.reserved[1] forces both).  Same results, same
    seeding work."""
    case = pu.make_case(seed=31, n_pairs=6000, genome_len=150000, contigs=2, n_rate=0.002, paired=1)
    ix = pu.build_index(case)
    compact = pu.cuda_results(case, ix)
    wide = pu.cuda_results(case, ix, reserved=(0, 1, 0, 0, 0))
    pu.assert_same(wide, compact)
    for k in ("seed_blocks", "sa_reads"):
        assert wide["stats"][k] == compact["stats"][k], k
    assert wide["stats"]["locate_blocks"] > compact["stats"]["locate_blocks"]      # 7 vs 3 steps per location on average
    pu.assert_same(wide, pu.oracle_results(case, ix))


def test_double_buffered_feed_equals_one_call(built):
    """mc_stage_batch_async(batch i+1) while mc_map_staged(batch i) runs: same records, totals and profile as mc_map_batch."""
    from mapcaller_b200 import api
    case = pu.make_case(seed=17, n_pairs=9000, genome_len=100000, n_rate=0.001)
    ix = pu.build_index(case)
    whole = pu.cuda_results(case, ix)
    seq, off = case["seq"], case["off"]
    n = len(off) - 1
    cuts = [0, 4000, 9000, 13000, n]
    parts = []
    for b, e in zip(cuts, cuts[1:]):
        ps = api.pinned_array((int(off[e] - off[b]),), np.uint8); ps[:] = seq[off[b]:off[e]]
        po = api.pinned_array((e - b + 1,), np.int64); po[:] = off[b:e + 1] - off[b]
        parts.append((ps, po))
    got = dict(reads=[], est=[])
    with api.Context(ix, want_alignments=1, update_profile=1, **case["params"]) as ctx:
        ctx.stage_batch_async(*parts[0], 0)
        for i in range(len(parts)):
            if i + 1 < len(parts):
                ctx.stage_batch_async(*parts[i + 1], (i + 1) & 1)
            res = ctx.map_staged(i & 1, copy=True)
            got["reads"] += api.unpack_reads(res); got["est"] += [int(x) for x in res["chunks"]["est_distance"]]
        t = ctx.totals()
        got["counters"] = dict(reads=t["total_reads"], mapped=t["total_mapped"], paired=t["total_paired"], dist_sum=t["total_distance"],
                               len_sum=t["read_length_sum"], avgDist=t["avg_dist"])
        got["profile"] = ctx.profile_columns(); got["ins"], got["dele"] = ctx.indels(); got["bp"] = ctx.breakpoints()
        got["inv"] = sorted(ctx.sites(0), key=lambda x: x[0]); got["tnl"] = sorted(ctx.sites(1), key=lambda x: x[0])
    pu.assert_same(got, whole)


def _fastq_text(reads, mate):
    import io
    from mapcaller_b200 import simulate as sim
    with tempfile.TemporaryDirectory() as td:
        sim.write_fastq(os.path.join(td, "x.fq"), reads, mate)
        return open(os.path.join(td, "x.fq"), "rb").read()


@pytest.mark.parametrize("layout", ["two_files", "interleaved", "unterminated"])
def test_fastq_ingest_on_device(built, layout):
    """mc_ingest_fastq (GetNextEntry / GetNextChunk, reference src/GetData.cpp:32-99): raw FASTQ blocks that end in the middle
    of records, fed block by block, give the same batches - hence the same records, totals and profile - as the host-parsed
    arrays; mates from two files or adjacent in one; a last line without newline is still a line."""
    from mapcaller_b200 import api
    case = pu.make_case(seed=23, n_pairs=2600, genome_len=60000, n_rate=0.002)
    ix = pu.build_index(case)
    whole = pu.cuda_results(case, ix)
    t1, t2 = _fastq_text(case["r1"], 1), _fastq_text(case["r2"], 2)
    if layout == "interleaved":
        a, b = t1.split(b"\n"), t2.split(b"\n")
        recs = []
        for i in range(0, len(a) - 1, 4):
            recs += a[i:i + 4] + b[i:i + 4]
        t1, t2 = b"\n".join(recs) + b"\n", None
    if layout == "unterminated":
        t1, t2 = t1[:-1], t2[:-1]
    got = dict(reads=[], est=[])
    with api.Context(ix, want_alignments=1, update_profile=1, **case["params"]) as ctx:
        p1 = p2 = 0
        blk1, blk2 = (150000, 157000) if t2 is not None else (310000, 0)
        for it in range(100):
            b1 = t1[p1:p1 + blk1]; b2 = None if t2 is None else t2[p2:p2 + blk2]
            final = p1 + blk1 >= len(t1) and (t2 is None or p2 + blk2 >= len(t2))
            r = ctx.ingest_fastq(b1, b2, slot=it & 1, final=final)
            assert final or (r["n_reads"] > 0 and r["n_reads"] % 200 == 0)
            assert r["consumed1"] <= len(b1)
            p1 += r["consumed1"]; p2 += r["consumed2"]
            res = ctx.map_staged(it & 1, copy=True)
            got["reads"] += api.unpack_reads(res); got["est"] += [int(x) for x in res["chunks"]["est_distance"]]
            if final:
                break
        assert p1 == len(t1)
        t = ctx.totals()
        got["counters"] = dict(reads=t["total_reads"], mapped=t["total_mapped"], paired=t["total_paired"], dist_sum=t["total_distance"],
                               len_sum=t["read_length_sum"], avgDist=t["avg_dist"])
        got["profile"] = ctx.profile_columns(); got["ins"], got["dele"] = ctx.indels(); got["bp"] = ctx.breakpoints()
        got["inv"] = sorted(ctx.sites(0), key=lambda x: x[0]); got["tnl"] = sorted(ctx.sites(1), key=lambda x: x[0])
        # malformed input: an empty read line
        with pytest.raises(api.McError):
            ctx.ingest_fastq(b"@x\n\n+\n\n@y\nACGT\n+\nIIII\n", None, slot=2, final=True)
    pu.assert_same(got, whole)


def test_pipelined_large_batch_equals_resident_path(built):
    """A profile-only batch of >= 400 k reads is cut into pieces whose upload overlaps the mapping of the previous piece
    (copy stream); the result must equal the single-piece path (mc_stage_batch + mc_map_staged) and the chunk-wise path."""
    from mapcaller_b200 import api
    case = pu.make_case(seed=16, n_pairs=210000, genome_len=400000, contigs=2)
    ix = pu.build_index(case)
    seq, off = case["seq"], case["off"]
    outs = []
    for mode in ("pipelined", "staged", "pinned_pipelined"):
        with api.Context(ix, paired=1, update_profile=1) as ctx:
            if mode == "pipelined":
                res = ctx.map_batch(seq, off)
            elif mode == "staged":
                ctx.stage_batch(seq, off, 0); res = ctx.map_staged(0, copy=True)
            else:
                ps = api.pinned_array(seq.shape, np.uint8); ps[:] = seq
                po = api.pinned_array(off.shape, np.int64); po[:] = off
                res = ctx.map_batch(ps, po)
            ins, dele = ctx.indels()
            outs.append(dict(chunks=res["chunks"].tolist(), totals=ctx.totals(), profile=ctx.profile_columns(), ins=ins, dele=dele, bp=ctx.breakpoints(),
                             inv=sorted(ctx.sites(0)), tnl=sorted(ctx.sites(1))))
    for o in outs[1:]:
        assert o["totals"] == outs[0]["totals"] and o["chunks"] == outs[0]["chunks"]
        assert np.array_equal(o["profile"], outs[0]["profile"])
        for k in ("ins", "dele", "bp", "inv", "tnl"):
            assert o[k] == outs[0][k], k
    assert outs[0]["totals"]["total_reads"] == 420000


def test_edge_cases(built):
    """Empty batch, reads shorter than a seed, all-N reads, unmappable reads, a read longer than the limit."""
    from mapcaller_b200 import api
    case = pu.make_case(seed=14, n_pairs=50, genome_len=30000)
    ix = pu.build_index(case)
    with api.Context(ix, paired=1, want_alignments=1) as ctx:
        res = ctx.map_batch(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64))
        assert len(res["chunks"]) == 0
        reads = [b"ACGT", b"ACGTACGTACGTACGT", b"N" * 100, b"ACGT" * 25, case["r1"][0].tobytes(), case["r2"][0].tobytes(), b"", b"A"]
        seq = np.frombuffer(b"".join(reads), dtype=np.uint8)
        off = np.zeros(len(reads) + 1, dtype=np.int64); off[1:] = np.cumsum([len(r) for r in reads])
        res = ctx.map_batch(seq, off)
        got = api.unpack_reads(res)
        assert [r["score"] for r in got[:4]] == [0, 0, 0, 0] and got[4]["score"] > 0 and got[5]["score"] > 0
        with pytest.raises(api.McError):
            ctx.map_batch(np.zeros(5000, dtype=np.uint8) + 65, np.array([0, 4500, 5000], dtype=np.int64))
        with pytest.raises(api.McError):
            ctx.map_batch(seq[:off[3]], off[:4])     # odd number of reads in paired mode
    case2 = dict(case, seq=seq, off=off)
    pu.assert_same(pu.cuda_results(case2, ix), pu.oracle_results(case2, ix))


def ragged_case(seed=19, n_pairs=6000, genome_len=120000):
    """Reads of very different lengths in one library (trimmed reads): 20 .. 150 bases, a few shorter than a seed."""
    case = pu.make_case(seed=seed, n_pairs=n_pairs, genome_len=genome_len, read_len=150, frag_mean=420, frag_sd=60, contigs=2)
    rng = np.random.default_rng(seed)
    seq, off = case["seq"], case["off"]
    n = len(off) - 1
    keep = rng.integers(20, 151, size=n); keep[rng.random(n) < 0.01] = rng.integers(1, 16, size=int((rng.random(n) < 0.01).sum()) or 1)[0]
    parts = [seq[off[i]:off[i] + min(int(keep[i]), int(off[i + 1] - off[i]))] for i in range(n)]
    noff = np.zeros(n + 1, dtype=np.int64); noff[1:] = np.cumsum([len(p) for p in parts])
    return dict(case, seq=np.concatenate(parts), off=noff)


def test_ragged_read_lengths(built):
    case = ragged_case()
    ix = pu.build_index(case)
    mine = pu.cuda_results(case, ix, batch_reads=5000)
    pu.assert_same(mine, pu.oracle_results(case, ix))


@pytest.mark.parametrize("wide", [0, 1])
def test_bwt_search_operator_matches_oracle(built, wide):
    """mc_bwt_search_batch (BWT_Search, reference src/bwt_search.cpp:121) on its own: match length, frequency and the set of
    locations for exact, mutated, repeat and N-interrupted queries, on both device index layouts."""
    import cpu_oracle
    from mapcaller_b200 import api, simulate as sim
    case = pu.make_case(seed=21, n_pairs=10, genome_len=60000, contigs=2, n_dup=12)
    ix = pu.build_index(case)
    ref = case["ref"]
    rnd = np.random.default_rng(5)
    qs, starts = [], []
    for i in range(1500):
        n = int(rnd.integers(17, 120)); p = int(rnd.integers(0, len(ref) - n))
        q = ref[p:p + n].copy()
        if i % 3 == 1:
            q = sim.revcomp(q)
        if i % 4 == 2:
            q[int(rnd.integers(0, n))] = ord("ACGT"[int(rnd.integers(0, 4))])
        if i % 7 == 3:
            q[int(rnd.integers(1, n))] = ord("N")
        code = np.array([{65: 0, 67: 1, 71: 2, 84: 3}.get(int(c) & 0xDF, 4) for c in q], dtype=np.uint8)
        st = int(rnd.integers(0, max(1, n - 16)))
        if code[st] > 3:
            st = 0 if code[0] <= 3 else 1
        qs.append(code); starts.append(st)
    codes = np.concatenate(qs); off = np.zeros(len(qs) + 1, dtype=np.int64); off[1:] = np.cumsum([len(q) for q in qs])
    with api.Context(ix, reserved=(0, wide, 0, 0, 0)) as ctx:
        ln, fr, loc = ctx.bwt_search_batch(codes, off, np.array(starts, dtype=np.int32))
    with tempfile.TemporaryDirectory() as td:
        ix.save(os.path.join(td, "idx"))
        orc = cpu_oracle.Oracle(os.path.join(td, "idx"), **case["params"])
        n_hits = 0
        for i, q in enumerate(qs):
            l, f, lo = orc.bwt_search(q, starts[i], len(q))
            assert (ln[i], fr[i]) == (l, f), (i, ln[i], fr[i], l, f)
            assert np.array_equal(loc[i], np.sort(lo)), i
            n_hits += f
        orc.close()
    assert n_hits > 1000


@pytest.mark.skipif(not pu.have_ref(), reason="oracle/_ref not on this box")
def test_seed_cluster_operator_matches_reference(built):
    """mc_seed_cluster_batch against the reference's own IdentifySimplePairs + SimplePairClustering (src/ReadMapping.cpp:125,194)
    read by read: the sorted simple pairs and the candidate clusters (scores, members, tandem-repeat resolution)."""
    import ref_oracle as ro
    from mapcaller_b200 import api
    case = pu.make_case(seed=9, n_pairs=1500, genome_len=60000, sv=5.0, n_dup=30, tandem=20, n_rate=0.004, contigs=2)
    ix = pu.build_index(case)
    seq, off = case["seq"], case["off"]
    n = len(off) - 1
    with api.Context(ix, paired=1) as ctx:
        got = ctx.seed_cluster_batch(seq, off)
    with tempfile.TemporaryDirectory() as td:
        ix.save(os.path.join(td, "idx")); ro.load(os.path.join(td, "idx"))
        n_clusters = 0
        for r in range(n):
            want = ro.seed_cluster(bytes(seq[off[r]:off[r + 1]]))
            assert got[r] == want, "read %d" % r
            n_clusters += len(want[1])
    assert n_clusters > n


def test_gapped_fill_kernel_matches_oracle(built):
    """mc_align_batch (the DP kernel on its own) vs the oracle's nw / ksw2 on random, mutated and tandem inputs."""
    from mapcaller_b200 import api
    rnd = random.Random(11)

    def rs(n, alpha="ACGT"):
        return "".join(rnd.choice(alpha) for _ in range(n))

    probs = []
    for _ in range(4000):
        m = rnd.randint(1, 90)
        k = rnd.random()
        if k < 0.3:
            a, b = rs(m), rs(rnd.randint(1, 90))
        elif k < 0.8:
            a = rs(m)
            b = "".join(c if rnd.random() > 0.15 else rnd.choice(["", "A", "C", "G", "T", "N", c + rnd.choice("ACGT")]) for c in a) or "A"
        else:
            u = rs(rnd.randint(1, 3))
            a, b = (u * 50)[:m], (u * 50)[:rnd.randint(1, 90)]
        probs.append((a.encode(), b.encode()))
    probs.append((b"A" * 300, b"A" * 250 + b"C" * 70))
    case = pu.make_case(seed=14, n_pairs=10, genome_len=20000)
    with api.Context(pu.build_index(case)) as ctx:
        for ksw2 in (False, True):
            mine = ctx.align_batch(probs, ksw2=ksw2)
            for p, m in zip(probs, mine):
                assert m == cpu_oracle.align(not ksw2, p[0], p[1]), (ksw2, p)


def test_smoke(built):
    pu.smoke_case()
