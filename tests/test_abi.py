"""The C-ABI library loads, exports every symbol include/mapcaller_b200.h declares, and has no CPU fallback."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mapcaller_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built):
    from mapcaller_b200 import api
    L = api.lib()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libmapcaller_b200.so does not export %s" % n
    assert b"sm_100a" in L.mc_version()


def test_header_compiles_as_plain_c(built, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mapcaller_b200.h"\nint main(void){ mc_params p; mc_params_default(&p);\n'
                   '  if (sizeof(mc_sam_rec) != 64 || sizeof(mc_variant_rec) != 48 || sizeof(mc_vc_params) != 32) return 2;   /* api.SAM_DT / VARIANT_DT / VcParams */\n'
                   '  return p.max_dup == 5 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    lib = os.path.join(ROOT, "mapcaller_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", lib, "-lmapcaller_b200", "-Wl,-rpath," + lib])
    assert subprocess.call([str(exe)]) == 0


def test_struct_layouts_match_ctypes(built):
    from mapcaller_b200 import api
    assert ctypes.sizeof(api.Params) == 16 * 4
    assert api.READ_DT.itemsize == 24 and api.CAND_DT.itemsize == 24 and api.FRAG_DT.itemsize == 40
    assert api.PAIR_DT.itemsize == 24 and api.CHUNK_DT.itemsize == 32 and api.PROFILE_DT.itemsize == 16


def test_argument_errors(built):
    from mapcaller_b200 import api
    L = api.lib()
    h = ctypes.c_void_p()
    assert L.mc_index_load(b"/nonexistent/prefix", ctypes.byref(h)) == -3
    assert b"cannot open" in L.mc_last_error()
    codes = np.zeros(10, dtype=np.uint8)
    lens = np.array([7], dtype=np.int32)
    assert L.mc_index_build(codes.ctypes.data, 10, 1, lens.ctypes.data, None, 1, ctypes.byref(h)) == -1


def _has_gpu() -> bool:
    return shutil.which("nvidia-smi") is not None and subprocess.call(["nvidia-smi", "-L"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(built):
    """Without a CUDA device the product refuses to map: there is no CPU path behind the ABI."""
    from mapcaller_b200 import api
    ix = api.Index.build(np.random.default_rng(0).integers(0, 4, 5000).astype(np.uint8))
    with pytest.raises(api.McError) as e:
        api.Context(ix)
    assert "(-2)" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """Nothing under mapcaller_b200/ or include/ may reference oracle/ (SPEC: the oracle is only a checker)."""
    for base in ("mapcaller_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "libmcoracle" not in txt and "libmcref" not in txt and "oracle/restate" not in txt, os.path.join(dp, f)
