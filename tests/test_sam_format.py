"""Host side of the SAM records: the line printer (api.format_sam_line) follows the reference's format strings
(src/SamReport.cpp:332,351,401,431-434) and GetComplementarySeq (src/tools.cpp:5-29).  No GPU involved."""
import numpy as np

from mapcaller_b200 import api


def rec(**kw):
    r = np.zeros(1, dtype=api.SAM_DT)[0]
    r["chrom"], r["nm"] = -1, -1
    for k, v in kw.items():
        r[k] = v
    return r


def test_complement_read():
    assert api.complement_read(b"ACGTNacgtX") == b"NacgtNACGT"
    assert api.complement_read(b"") == b""


def test_unmapped_line():
    l = api.format_sam_line(rec(flag=77), b"*", b"r1", b"ACGT", b"IIII", [b"chr1"])
    assert l == b"r1\t77\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\tAS:i:0\tXS:i:0"
    l = api.format_sam_line(rec(flag=141, reverse=1), b"*", b"r1", b"AACG", b"ABCD", [b"chr1"])   # an unmapped mate 2 is printed reversed
    assert l == b"r1\t141\t*\t0\t0\t*\t*\t0\t0\tCGTT\tDCBA\tAS:i:0\tXS:i:0"


def test_mapped_lines():
    r = rec(flag=99, chrom=1, pos=1234, mapq=60, nm=2, xs=0, has_mate=1, mate_pos=1500, tlen=366)
    r["as"] = 98
    l = api.format_sam_line(r, b"40M2D60M", b"q", b"ACGT", b"IIII", [b"chr1", b"chr2"])
    assert l == b"q\t99\tchr2\t1234\t60\t40M2D60M\t=\t1500\t366\tACGT\tIIII\tNM:i:2\tAS:i:98\tXS:i:0"
    r = rec(flag=16, chrom=0, pos=7, mapq=0, nm=0, xs=100, reverse=1)
    r["as"] = 100
    l = api.format_sam_line(r, b"4M", b"q", b"AACG", None, [b"chr1"])   # FASTA reads: no qualities
    assert l == b"q\t16\tchr1\t7\t0\t4M\t*\t0\t0\tCGTT\t*\tNM:i:0\tAS:i:100\tXS:i:100"
    assert api.format_sam_line(rec(flag=-1), b"*", b"q", b"A", b"I", [b"c"]) is None   # the reference prints nothing for such a read


def test_golden_sam_fixture_is_the_references_output(tmp_path):
    """tests/golden/sam_*.golden pinned against the reference CLI where oracle/_ref is on the box."""
    import os
    import pytest
    import golden_util as gu
    import parity_util as pu
    if not os.path.exists(os.path.join(pu.ROOT, "oracle", "_ref", "MapCaller")):
        pytest.skip("oracle/_ref not on this box")
    for name in ("pe_nw", "se_nw"):
        case = gu.with_mates(gu.load(name)[0])
        d = tmp_path / name
        d.mkdir()
        ref = pu.sam_comparable(pu.sam_lines_reference(case, str(d)), bool(case["params"]["paired"]))
        assert ref == gu.load_sam(name)
        assert len(ref) >= 800
