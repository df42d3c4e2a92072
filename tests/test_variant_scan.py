"""Variant-calling scan, CPU side: the committed fixtures tests/golden/vc_*.golden are what the unmodified reference
(oracle/_ref: CalBlockReadDepth + IdentifyVariants, src/VariantCalling.cpp:106-120,550-680) produces today, and the
C-ABI entry refuses bad arguments.  The CUDA path is compared with the same fixtures in tests/test_variant_scan_gpu.py."""
import ctypes

import numpy as np
import pytest

import golden_util as gu
import parity_util as pu
import ref_oracle as ro

VC_GOLDEN = ("pe_nw", "pe_multi")


@pytest.mark.skipif(not pu.have_ref(), reason="oracle/_ref not on this box")
@pytest.mark.parametrize("name", VC_GOLDEN)
def test_fixture_is_the_references_output(built, name):
    case, _ = gu.load(name)
    sets, gold = gu.load_vc(name)
    assert sets == pu.VC_SETS
    ref = pu.ref_results(case, pu.build_index(case), want_reads=False, vc=sets)
    for (gv, gd), (rv, rd) in zip(gold, ref["vc"]):
        assert np.array_equal(gd, rd)
        assert ro.variants_equal(gv, rv)


def test_fixture_shape():
    for name in VC_GOLDEN:
        sets, gold = gu.load_vc(name)
        assert len(sets) == len(gold) == len(pu.VC_SETS)
        types = {v["VarType"] for vs, _ in gold for v in vs}
        assert {0, 1, 2, 5, 6, 10, 11} <= types, types   # SUB INS DEL CNV UMR NOR MON all occur
        for vs, depth in gold:
            assert [(v["gPos"], v["VarType"]) for v in vs] == sorted((v["gPos"], v["VarType"]) for v in vs)
            assert depth.dtype == np.int32 and len(depth) > 0


def test_abi_layout_and_argument_errors(built):
    from mapcaller_b200 import api
    L = api.lib()
    assert ctypes.sizeof(api.VcParams) == 32 and api.VARIANT_DT.itemsize == 48
    vp = api.VcParams()
    L.mc_vc_params_default(ctypes.byref(vp))
    assert (vp.min_allele_depth, vp.ploidy, vp.min_cnv_size, vp.min_unmapped_size, vp.gvcf, vp.monomorphic, vp.somatic) == (5, 2, 50, 50, 0, 0, 0)
    assert abs(vp.frequency_thr - 0.2) < 1e-7
    r, n, a, d, nb = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
    rc = L.mc_variant_scan(None, ctypes.byref(vp), ctypes.byref(r), ctypes.byref(n), ctypes.byref(a), ctypes.byref(d), ctypes.byref(nb))
    assert rc == -1 and b"null" in L.mc_last_error()
