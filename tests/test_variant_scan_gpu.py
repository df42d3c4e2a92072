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
