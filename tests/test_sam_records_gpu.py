"""mc_sam_records (flag, RNAME / POS, MAPQ, CIGAR, mate fields, NM / AS / XS computed on the device) against the SAM text of
the unmodified reference CLI (oracle/_ref/MapCaller -t 1): the lines printed from the device records plus the FASTQ text
must equal the reference's lines byte for byte (single-end: up to QUAL, see parity_util.sam_comparable)."""
import os

import pytest

import golden_util as gu
import parity_util as pu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "MapCaller")

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not on this box")

CASES = {
    "pe_two_contigs_sv": dict(seed=41, n_pairs=6000, genome_len=120000, contigs=2, sv=3.0),
    "se_sv": dict(seed=42, n_pairs=4000, genome_len=80000, paired=0, sv=3.0),
    "pe_ksw2_indels_n": dict(seed=43, n_pairs=4000, genome_len=90000, contigs=3, alg_ksw2=1, indel_rate=0.003, n_rate=0.004),
    "pe_repeats": dict(seed=44, n_pairs=8000, genome_len=100000, sv=5.0, n_dup=30, tandem=20),
    "pe_long_250": dict(seed=45, n_pairs=3000, genome_len=150000, read_len=250, frag_mean=600, frag_sd=80, indel_rate=0.003, contigs=2),
    "pe_params": dict(seed=47, n_pairs=3000, genome_len=60000, max_clip=2, max_dup=15, max_mismatch_rate=0.1, contigs=4),
    "se_diverged_repeats": dict(seed=51, n_pairs=6000, genome_len=120000, repeat_frac=0.5, repeat_div=0.01, paired=0),
    "se_ksw2": dict(seed=48, n_pairs=3000, genome_len=80000, paired=0, alg_ksw2=1, indel_rate=0.004, contigs=2),
}


@pytest.mark.parametrize("name", ("pe_nw", "pe_ksw2", "se_nw", "pe_multi"))
def test_sam_lines_equal_the_golden_fixture(built, name):
    """Committed output of the unmodified reference CLI (tests/golden/make_golden_sam.py)."""
    case = gu.with_mates(gu.load(name)[0])
    mine = pu.sam_comparable(pu.sam_lines_cuda(case), bool(case["params"]["paired"]))
    assert mine == gu.load_sam(name)


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_sam_lines_equal_the_reference_cli(built, tmp_path, name):
    case = pu.make_case(**CASES[name])
    paired = bool(case["params"]["paired"])
    # two batches: the records always describe the batch mapped last
    mine = pu.sam_comparable(pu.sam_lines_cuda(case, batch_reads=4000 if name == "pe_repeats" else None), paired)
    ref = pu.sam_comparable(pu.sam_lines_reference(case, str(tmp_path)), paired)
    assert len(mine) == len(ref) > 1000
    bad = [k for k, (a, b) in enumerate(zip(mine, ref)) if a != b]
    assert not bad, "%d SAM lines differ, first:\n%r\n%r" % (len(bad), mine[bad[0]], ref[bad[0]])
    assert any(b"S" in l.split(b"\t")[5] for l in ref) and any(l.split(b"\t")[2] == b"*" for l in ref)


def test_sam_records_need_a_mapped_batch(built):
    from mapcaller_b200 import api
    case = pu.make_case(seed=3, n_pairs=200, genome_len=20000)
    with api.Context(pu.build_index(case), **case["params"]) as ctx:
        with pytest.raises(api.McError):
            ctx.sam_records()


def test_sam_record_properties_at_ecoli_size(built):
    """configs[1] genome size: properties of the records that do not depend on the size (CIGAR spans the read, positions lie
    inside their chromosome, flag bits agree with the record, mates that name each other have opposite TLEN)."""
    import re
    import numpy as np
    from mapcaller_b200 import api, simulate as sim
    G, P, L = 4_600_000, 50_000, 100
    parts = [sim.genome(G // 2, 7, n_dup=100), sim.genome(G - G // 2, 8, n_dup=100)]
    g = np.concatenate(parts)
    mut, _ = sim.mutate(g, 9, snp_per_mb=3000, small_indel_per_mb=200, large_indel_per_mb=50, sv_per_mb=1)
    r1, r2 = sim.simulate_pairs(mut, P, L, seed=11)
    r1[::97, 10:40] = ord("N")   # some reads that cannot be placed
    seq, off = sim.interleave(r1, r2)
    with api.Context(api.Index.build(sim.encode(g), [len(x) for x in parts], ["a", "b"]), paired=1, update_profile=0) as ctx:
        ctx.map_batch(seq, off)
        recs, cigars = ctx.sam_records()
    assert len(recs) == 2 * P
    mapped = recs["chrom"] >= 0
    assert mapped.mean() > 0.9 and (~mapped).sum() > 0
    assert ((recs["flag"] & 4 != 0) == ~mapped).all() and (recs["flag"] & 1 == 1).all()
    assert (recs["flag"][0::2] & 0x40 != 0).all() and (recs["flag"][1::2] & 0x80 != 0).all()
    assert (recs["mapq"] >= 0).all() and (recs["mapq"] <= 60).all() and (recs["as"] >= recs["xs"]).all()
    assert (recs["nm"][mapped] >= 0).all() and (recs["nm"][mapped] + recs["as"][mapped] == L).all() and (recs["nm"][~mapped] == -1).all()
    clen = np.array([len(parts[c]) if c >= 0 else 0 for c in recs["chrom"]])
    assert (recs["pos"][mapped] >= 1).all() and (recs["pos"][mapped] <= clen[mapped]).all()
    for k in np.nonzero(mapped)[0][:20000]:
        ops = re.findall(rb"(\d+)([MIDS])", cigars[k])
        assert b"".join(a + b for a, b in ops) == cigars[k]
        assert sum(int(a) for a, b in ops if b in b"MIS") == L, cigars[k]
    a, b = recs[0::2], recs[1::2]
    both = (a["has_mate"] == 1) & (b["has_mate"] == 1) & (a["mate_pos"] == b["pos"]) & (b["mate_pos"] == a["pos"])
    assert both.mean() > 0.8 and (a["tlen"][both] == -b["tlen"][both]).all()


# ---- mc_sam_text: the whole SAM line assembled on the device from the FASTQ text mc_ingest_fastq left in HBM -------------
@pytest.mark.parametrize("name", ("pe_nw", "pe_ksw2", "se_nw", "pe_multi"))
def test_device_sam_text_equals_the_golden_fixture(built, name):
    case = gu.with_mates(gu.load(name)[0])
    mine = pu.sam_comparable(pu.sam_text_cuda(case), bool(case["params"]["paired"]))
    assert mine == gu.load_sam(name)


TEXT_CASES = dict(CASES, pe_lower_case_n=dict(seed=15, n_pairs=3000, genome_len=60000, lower_rate=0.3, n_rate=0.01),
                  pe_all_best=dict(seed=9, n_pairs=4000, genome_len=60000, n_dup=40, tandem=10),
                  se_all_best=dict(seed=10, n_pairs=4000, genome_len=60000, n_dup=40, tandem=10, paired=0))


@needs_ref
@pytest.mark.parametrize("name", sorted(TEXT_CASES))
def test_device_sam_text_equals_the_reference_cli(built, tmp_path, name):
    """QNAME / SEQ / QUAL come from the FASTQ text in the slot (reverse-complemented / reversed on the reverse strand, mate 2
    folded as ReverseOrientation + GetComplementarySeq do); *_all_best = the reference's -m (one line per best candidate)."""
    case = pu.make_case(**TEXT_CASES[name])
    paired = bool(case["params"]["paired"]); all_best = name.endswith("all_best")
    mine = pu.sam_comparable(pu.sam_text_cuda(case, batch_pairs=2000 if name == "pe_repeats" else None, all_best=all_best), paired)
    ref = pu.sam_comparable(pu.sam_lines_reference(case, str(tmp_path), all_best=all_best), paired)
    if all_best and not paired:
        # single-end -m: SetSingledAlignmentFlag (src/SamReport.cpp:7-24) only initialises the flag of the first best candidate
        # and the lines of the others print an uninitialised SamFlag - every column but FLAG is compared
        strip = lambda ls: [b"\t".join(l.split(b"\t")[:1] + l.split(b"\t")[2:]) for l in ls]
        mine, ref = strip(mine), strip(ref)
    assert len(mine) == len(ref) > 1000
    bad = [k for k, (a, b) in enumerate(zip(mine, ref)) if a != b]
    assert not bad, "%d SAM lines differ, first:\n%r\n%r" % (len(bad), mine[bad[0]], ref[bad[0]])
    if all_best:
        assert len(ref) > (2 if paired else 1) * len(case["r1"])      # some reads have several best candidates


def test_device_sam_text_needs_the_fastq_text(built):
    from mapcaller_b200 import api
    case = pu.make_case(seed=3, n_pairs=200, genome_len=20000)
    with api.Context(pu.build_index(case), **case["params"]) as ctx:
        ctx.stage_batch(case["seq"], case["off"], 0); ctx.map_staged(0)
        with pytest.raises(api.McError):
            ctx.sam_text(0)
