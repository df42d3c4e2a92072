"""Pins the CPU restatement (oracle/restate) against the committed golden vectors, which are outputs of the
unmodified reference, and - where oracle/_ref exists - against the reference itself on further seeded cases."""
import random

import pytest

import cpu_oracle
import golden_util as gu
import parity_util as pu
import ref_oracle


@pytest.mark.parametrize("name", sorted(gu.GOLDEN_CASES))
def test_restatement_matches_golden(built, name):
    case, ref = gu.load(name)
    ix = pu.build_index(case)
    mine = pu.oracle_results(case, ix)
    pu.assert_same(mine, ref, paired=bool(case["params"]["paired"]))


@pytest.mark.skipif(not pu.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kw", [dict(seed=21, n_pairs=4000, genome_len=120000, contigs=2, sv=3.0, n_dup=20, tandem=10),
                                dict(seed=22, n_pairs=2500, genome_len=60000, alg_ksw2=1, indel_rate=0.004, read_len=150, frag_mean=450),
                                dict(seed=23, n_pairs=3000, genome_len=50000, n_rate=0.01, sub_rate=0.03, max_dup=2),
                                dict(seed=24, n_pairs=2000, genome_len=50000, paired=0, max_clip=2, max_pos_diff=10)])
def test_restatement_matches_reference(built, kw):
    case = pu.make_case(**kw)
    ix = pu.build_index(case)
    pu.assert_same(pu.oracle_results(case, ix), pu.ref_results(case, ix), paired=bool(case["params"]["paired"]))


@pytest.mark.skipif(not ref_oracle.available(), reason="oracle/_ref not built")
def test_gapped_fills_match_reference(built):
    """nw_alignment / ksw2_alignment of the reference vs the restated integer DPs on random, mutated and tandem inputs."""
    rnd = random.Random(7)

    def rs(n, alpha="ACGT"):
        return "".join(rnd.choice(alpha) for _ in range(n))

    probs = []
    for _ in range(1500):
        m = rnd.randint(1, 70)
        k = rnd.random()
        if k < 0.3:
            a, b = rs(m), rs(rnd.randint(1, 70))
        elif k < 0.8:
            a = rs(m)
            b = "".join(c if rnd.random() > 0.15 else rnd.choice(["", "A", "C", "G", "T", "N", c + rnd.choice("ACGT")]) for c in a) or "A"
        else:
            u = rs(rnd.randint(1, 3))
            a, b = (u * 40)[:m], (u * 40)[:rnd.randint(1, 70)]
        probs.append((a.encode(), b.encode()))
    for use_nw in (True, False):
        for a, b in probs:
            assert cpu_oracle.align(use_nw, a, b) == ref_oracle.align(use_nw, a, b), (use_nw, a, b)
