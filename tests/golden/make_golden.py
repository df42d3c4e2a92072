"""Generates tests/golden/*.golden from the unmodified reference (oracle/_ref).  Run in the build container:
    python tests/golden/make_golden.py
The index is built by the product's own builder (mc_index_build), which tests/test_index.py pins byte-for-byte
against the reference's builder, so only host code of the library is involved here - no GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402
import parity_util as pu  # noqa: E402

if __name__ == "__main__":
    assert pu.have_ref(), "oracle/_ref is missing: run `make -C oracle ref`"
    for name, kw in gu.GOLDEN_CASES.items():
        case = pu.make_case(**kw)
        ix = pu.build_index(case)
        ref = pu.ref_results(case, ix)
        gu.save(name, case, ref)
        print(name, os.path.getsize(gu.path(name)), "bytes;", len(ref["reads"]), "reads,", ref["counters"])
