"""Developer harness: the drop-in driver linked against the host-compiled stage bodies (tools/hostemu/build.sh) instead of the
CUDA library, so that the shim's host logic (block reader, libraries, SAM / VCF hand-over) can be debugged against the
reference CLI on a box without a GPU.  Runs the cases of tests/test_dropin_gpu.py.  Not a test, not a fallback.
usage: python tests/devharness/dropin_host.py [case ...]"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_dropin_gpu as td

OBJ = os.path.join(ROOT, "oracle", "_ref", "obj")
KEEP = "main GetData VariantCalling SamReport tools bwt_index bwa_utils bwa_bwt bwa_bntseq bwa_QSufSort bwa_bwt_gen bwa_bwtindex htslib_stubs".split()
HOST_BIN = os.path.join(ROOT, "tools", "hostemu", "_build", "MapCaller_hostemu")


def build():
    emu = os.path.join(ROOT, "tools", "hostemu", "_build")
    src = os.path.join(ROOT, "mapcaller_b200", "dropin", "ReadMapping_b200.cpp")
    obj = os.path.join(emu, "ReadMapping_b200.o")
    subprocess.check_call(["g++", "-O2", "-w", "-fPIC", "-I/root/reference/src", "-I" + os.path.join(ROOT, "include"), "-c", src, "-o", obj])
    weak = os.path.join(emu, "VariantCalling_weak.o")      # as mapcaller_b200/dropin/Makefile: IdentifyVariants comes from the shim
    subprocess.check_call(["objcopy", "--weaken-symbol=_Z16IdentifyVariantsPv", os.path.join(OBJ, "VariantCalling.o"), weak])
    subprocess.check_call(["g++", "-O2", obj, weak] + [os.path.join(OBJ, k + ".o") for k in KEEP if k != "VariantCalling"] + ["-o", HOST_BIN, os.path.join(emu, "libmc_hostemu.so"),
                           "-Wl,-rpath," + emu, "-lz", "-lm", "-lpthread"])


if __name__ == "__main__":
    build()
    td.GPU_BIN = HOST_BIN
    names = sys.argv[1:]
    import inspect, pathlib
    for fn_name, fn in inspect.getmembers(td, inspect.isfunction):
        if not fn_name.startswith("test_"):
            continue
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        params = marks[0].args[1] if marks else [None]
        for prm in params:
            tag = fn_name + ("[%s]" % prm if prm is not None else "")
            if names and not any(n in tag for n in names):
                continue
            with tempfile.TemporaryDirectory() as d:
                try:
                    fn(pathlib.Path(d), prm) if prm is not None else fn(pathlib.Path(d))
                    print("OK  ", tag, flush=True)
                except AssertionError as e:
                    print("FAIL", tag, str(e)[:300], flush=True)
