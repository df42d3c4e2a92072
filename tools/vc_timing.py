"""Timing of the device variant-calling scan (mc_variant_scan) on an E. coli-sized profile, beside the two things it
replaces: downloading the 16-byte-per-column profile (mc_profile_read) and the reference's own CalBlockReadDepth +
IdentifyVariants on the host (oracle/_ref, one thread - the reference forces iThreadNum = 1 there,
src/VariantCalling.cpp:717).  Results are compared record by record before anything is printed.
usage: python tools/vc_timing.py [genome_bp] [pairs]      (prints one JSON line)"""
import ctypes as C, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mapcaller_b200 import api, simulate as sim
if os.environ.get("MC_HOSTEMU_DEV"): api._LIB_PATH = os.path.join(ROOT, "tools", "hostemu", "_build", "libmc_hostemu.so")

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4_600_000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
g = sim.genome(G, 7, n_dup=200)
mut, _ = sim.mutate(g, 8, snp_per_mb=3000, small_indel_per_mb=200, large_indel_per_mb=50, sv_per_mb=1)
r1, r2 = sim.simulate_pairs(mut, P, 100, seed=11)
seq, off = sim.interleave(r1, r2)
ix = api.Index.build(sim.encode(g))
out = dict(genome_bp=G, pairs=P)
with api.Context(ix, paired=1) as ctx:
    ctx.map_batch(seq, off)
    L = api.lib(); vp = api.VcParams(); L.mc_vc_params_default(C.byref(vp))
    for mode in ("default", "gvcf"):
        vp.gvcf = int(mode == "gvcf")
        r, n, a, d, nb = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_void_p(), C.c_int64()
        ts = []
        for _ in range(5):
            t = time.perf_counter()
            rc = L.mc_variant_scan(ctx._h, C.byref(vp), C.byref(r), C.byref(n), C.byref(a), C.byref(d), C.byref(nb))
            ts.append(time.perf_counter() - t); assert rc == 0
        out["scan_ms_" + mode] = round(1000 * min(ts), 3); out["records_" + mode] = n.value
    ts = []
    for _ in range(3):
        t = time.perf_counter(); prof = ctx.profile(); ts.append(time.perf_counter() - t)
    out["profile_download_ms"] = round(1000 * min(ts), 3); out["profile_bytes"] = int(prof.nbytes)
    mine = [ctx.variant_scan(), ctx.variant_scan(gvcf=1)]
import ref_oracle as ro
if ro.available():
    with tempfile.TemporaryDirectory() as td:
        ix.save(os.path.join(td, "idx")); ro.load(os.path.join(td, "idx")); ro.set_params()
        t = time.perf_counter(); ro.map_reads(seq, off, True); out["ref_mapping_s_1thread"] = round(time.perf_counter() - t, 2)
        for mode, m in zip(("default", "gvcf"), mine):
            t = time.perf_counter(); rv, rd = ro.variant_scan(gvcf=int(mode == "gvcf")); out["ref_scan_ms_" + mode] = round(1000 * (time.perf_counter() - t), 1)
            assert np.array_equal(m[1], rd) and ro.variants_equal(m[0], rv), mode
        out["identical_to_reference"] = True
print(json.dumps(out))
