"""mc_variant_scan (device scan of the resident profile) against the reference's IdentifyVariants: the committed
fixtures of the unmodified reference (tests/golden/vc_*.golden) and, where oracle/_ref is on the box, live runs on cases
with long gap runs, repeats (dup runs), indel-rich reads and deep duplicates - defaults, gVCF, monomorphic, somatic,
haploid / low-threshold parameter sets each."""
import numpy as np
import pytest

import golden_util as gu
import parity_util as pu
import ref_oracle as ro

pytestmark = pytest.mark.gpu

LIVE = {
    "sparse_long_gaps": lambda: pu.sparse_case(),
    "sv_repeats": lambda: pu.make_case(seed=9, n_pairs=20000, genome_len=200000, sv=5.0, n_dup=30, tandem=20),
    "ksw2_indels": lambda: pu.make_case(seed=6, n_pairs=5000, genome_len=100000, alg_ksw2=1, indel_rate=0.002),
    "deep_duplicates": lambda: pu.make_case(seed=11, n_pairs=30000, genome_len=20000, max_dup=3),
}


@pytest.mark.parametrize("name", ("pe_nw", "pe_multi"))
def test_variant_scan_matches_golden(built, name):
    case, _ = gu.load(name)
    sets, gold = gu.load_vc(name)
    mine = pu.cuda_results(case, pu.build_index(case), want_reads=False, vc=sets)
    pu.assert_same_variants(mine, dict(vc=gold))
    # every SUB / INS / DEL / NOR / MON record carries the MappingRecord_t of its column
    prof = mine["profile"]
    for v in mine["vc"][2][0]:
        if v["VarType"] in (0, 1, 2, 11):
            w = v["record"][0]
            assert [(w >> s) & 0xFFF for s in (0, 12, 24, 36, 48)] == list(prof[v["gPos"], :5])


@pytest.mark.skipif(not pu.have_ref(), reason="oracle/_ref not on this box")
@pytest.mark.parametrize("name", sorted(LIVE))
def test_variant_scan_matches_reference(built, name):
    case = LIVE[name]()
    ix = pu.build_index(case)
    mine = pu.cuda_results(case, ix, want_reads=False, vc=pu.VC_SETS)
    ref = pu.ref_results(case, ix, want_reads=False, vc=pu.VC_SETS)
    pu.assert_same_variants(mine, ref)
    if name == "sparse_long_gaps":   # the interior gap is longer than 65535 columns: DP is its length modulo 2^16
        gaps = [v for v in ref["vc"][0][0] if v["VarType"] == 6]
        assert len(gaps) >= 2


def test_variant_scan_properties_at_ecoli_size(built):
    """configs[1] genome size (4.6 Mbp), where the reference's own scan is too slow for a unit test: properties that do not
    depend on the size - BlockDepthArr recomputed from the downloaded profile, (gPos, VarType) order, idempotence, every
    record consistent with its own MappingRecord_t and with the thresholds of src/VariantCalling.cpp:566-600."""
    import math
    from mapcaller_b200 import api, simulate as sim
    G, P = 4_600_000, 60_000
    g = sim.genome(G, 7, n_dup=200)
    mut, _ = sim.mutate(g, 8, snp_per_mb=3000, small_indel_per_mb=200, large_indel_per_mb=50, sv_per_mb=1)
    r1, r2 = sim.simulate_pairs(mut[:1_500_000], P, 100, seed=11)   # 8x over the first third, nothing elsewhere
    seq, off = sim.interleave(r1, r2)
    with api.Context(api.Index.build(sim.encode(g)), paired=1, update_profile=1) as ctx:
        ctx.map_batch(seq, off)
        recs, depth = ctx.variant_scan()
        again, depth2 = ctx.variant_scan()
        prof = ctx.profile_columns()
    assert recs == again and np.array_equal(depth, depth2)
    cov = prof[:, :4].sum(axis=1).astype(np.int64)
    pad = np.zeros((-len(cov)) % 100, dtype=np.int64)
    want = np.concatenate([cov, pad]).reshape(-1, 100).sum(axis=1) // 100
    assert np.array_equal(depth, want.astype(np.int32))
    keys = [(v["gPos"], v["VarType"]) for v in recs]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    kinds = {v["VarType"] for v in recs}
    assert {0, 1, 2, 6} <= kinds, kinds
    ref_code = sim.encode(g)
    for v in recs:
        p, t = v["gPos"], v["VarType"]
        if t in (0, 1, 2):
            w = v["record"][0]
            assert [(w >> s) & 0xFFF for s in (0, 12, 24, 36, 48)] == list(prof[p, :5])
        if t == 0:   # substitution: depth and allele thresholds, ALT is not the reference base
            c = int(cov[p]); thr = max(int(depth[p // 100]) >> 1, 5)
            assert v["DP"] == c >= thr and v["AD_ref"] == prof[p, ref_code[p]]
            alts = v["alt"].rstrip(b"\0").split(b",")
            assert all(prof[p, b"ACGT".index(a)] >= max(math.ceil(c * float(np.float32(0.2))), 5) and b"ACGT".index(a) != ref_code[p] for a in alts)
            assert v["AD_alt"] == sum(int(prof[p, b"ACGT".index(a)]) for a in alts)
        if t == 6:   # unmapped run: empty columns only, at least MinUnmappedSize of them (length modulo 2^16 in DP)
            assert cov[p] == 0 and prof[p, 4] == 0 and (p == 0 or cov[p - 1] > 0 or prof[p - 1, 4] > 0)
    gaps = [v for v in recs if v["VarType"] == 6]
    assert all(v["DP"] >= 50 or v["DP"] < 50 and cov[v["gPos"]:v["gPos"] + 65536].sum() == 0 for v in gaps)
