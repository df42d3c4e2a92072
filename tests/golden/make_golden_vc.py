"""Generates tests/golden/vc_*.golden: VariantVec / BlockDepthArr of the unmodified reference (oracle/_ref:
CalBlockReadDepth + IdentifyVariants + RemoveConsecutiveGenomicVariant, src/VariantCalling.cpp:106-120,550-694) on the
profiles of the committed golden cases, for every parameter set of parity_util.VC_SETS.  Run in the build container:
    python tests/golden/make_golden_vc.py"""
import os
import pickle
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402
import parity_util as pu  # noqa: E402

VC_GOLDEN = ("pe_nw", "pe_multi")

if __name__ == "__main__":
    assert pu.have_ref(), "oracle/_ref is missing: run `make -C oracle ref`"
    for name in VC_GOLDEN:
        case, _ = gu.load(name)
        ix = pu.build_index(case)
        ref = pu.ref_results(case, ix, want_reads=False, vc=pu.VC_SETS)
        payload = dict(sets=pu.VC_SETS, vc=[(v, d.tolist()) for v, d in ref["vc"]])
        with open(gu.path("vc_" + name), "wb") as fh:
            fh.write(zlib.compress(pickle.dumps(payload, protocol=4), 9))
        print(name, os.path.getsize(gu.path("vc_" + name)), "bytes;", [len(v) for v, _ in ref["vc"]], "variants")
