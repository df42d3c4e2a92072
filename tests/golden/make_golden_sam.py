"""Generates tests/golden/sam_*.golden: the SAM records the unmodified reference CLI (oracle/_ref/MapCaller -t 1) prints for
the committed golden cases (single-end: the columns before QUAL, see parity_util.sam_comparable).  Run in the build container:
    python tests/golden/make_golden_sam.py"""
import os
import sys
import tempfile
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402
import parity_util as pu  # noqa: E402

SAM_GOLDEN = ("pe_nw", "pe_ksw2", "se_nw", "pe_multi")

if __name__ == "__main__":
    assert pu.have_ref(), "oracle/_ref is missing: run `make -C oracle ref`"
    for name in SAM_GOLDEN:
        case = gu.with_mates(gu.load(name)[0])
        with tempfile.TemporaryDirectory() as td:
            lines = pu.sam_comparable(pu.sam_lines_reference(case, td), bool(case["params"]["paired"]))
        with open(gu.path("sam_" + name), "wb") as fh:
            fh.write(zlib.compress(b"\n".join(lines), 9))
        print(name, len(lines), "lines,", os.path.getsize(gu.path("sam_" + name)), "bytes")
