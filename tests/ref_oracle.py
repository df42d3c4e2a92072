"""ctypes access to oracle/_ref/libmcref.so (the UNMODIFIED reference compiled by oracle/Makefile).

Test infrastructure only.  The library keeps the reference's process-wide globals, so one Python
process can hold exactly one loaded index; tests that need several run helpers in subprocesses
(see `run_in_subprocess`).
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libmcref.so")
BIN_PATH = os.path.join(ROOT, "oracle", "_ref", "MapCaller")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.mcref_load.argtypes = [C.c_char_p]
        L.mcref_build_index.argtypes = [C.c_char_p, C.c_char_p]
        L.mcref_set_params.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.mcref_set_vc_flags.argtypes = [C.c_int] * 4
        L.mcref_genome_size.restype = C.c_int64
        L.mcref_refseq.restype = C.c_void_p
        L.mcref_bwt_words.restype = C.c_void_p
        L.mcref_bwt_words.argtypes = [C.POINTER(C.c_int64)]
        L.mcref_sa.restype = C.c_void_p
        L.mcref_sa.argtypes = [C.POINTER(C.c_int64)]
        L.mcref_bwt_meta.argtypes = [C.POINTER(C.c_uint64)]
        L.mcref_bwt_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        L.mcref_bwt_sa.restype = C.c_uint64
        L.mcref_bwt_sa.argtypes = [C.c_uint64]
        L.mcref_align.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]
        L.mcref_seed_cluster.restype = C.c_void_p
        L.mcref_seed_cluster.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
        L.mcref_map.restype = C.c_void_p
        L.mcref_map.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.mcref_run_mapping.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.mcref_variant_calling.argtypes = [C.c_char_p]
        L.mcref_variant_scan.restype = C.c_void_p
        L.mcref_variant_scan.argtypes = [C.c_int, C.c_float] + [C.c_int] * 6 + [C.POINTER(C.c_int64)]
        L.mcref_counters.argtypes = [C.POINTER(C.c_int64)]
        L.mcref_profile.argtypes = [C.c_int64, C.c_int64, C.c_void_p]
        L.mcref_indels.restype = C.c_void_p
        L.mcref_indels.argtypes = [C.c_int, C.POINTER(C.c_int64)]
        L.mcref_breakpoints.restype = C.c_void_p
        L.mcref_breakpoints.argtypes = [C.POINTER(C.c_int64)]
        L.mcref_sites.restype = C.c_void_p
        L.mcref_sites.argtypes = [C.c_int, C.POINTER(C.c_int64)]
        L.mcref_free.argtypes = [C.c_void_p]
        L.mcref_chrom.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        _lib = L
    return _lib


def _take(ptr, n) -> bytes:
    b = C.string_at(ptr, n.value)
    lib().mcref_free(ptr)
    return b


def build_index(fa: str, prefix: str) -> None:
    lib().mcref_build_index(fa.encode(), prefix.encode())


def load(prefix: str) -> int:
    rc = lib().mcref_load(prefix.encode())
    if rc != 0:
        raise RuntimeError("mcref_load failed: %d" % rc)
    return lib().mcref_genome_size()


def set_params(max_pos_diff=30, max_clip=5, max_dup=5, maxmm=0.05, nw=True, unique=True, threads=1):
    lib().mcref_set_params(max_pos_diff, max_clip, max_dup, maxmm, int(nw), int(unique), threads)


def index_arrays():
    """(bwt words uint32, sa uint64, meta dict) exactly as the reference holds them in memory."""
    n = C.c_int64()
    p = lib().mcref_bwt_words(C.byref(n))
    bwt = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n.value,)).copy()
    p = lib().mcref_sa(C.byref(n))
    sa = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n.value,)).copy()
    m = (C.c_uint64 * 8)()
    lib().mcref_bwt_meta(m)
    meta = dict(primary=m[0], L2=[m[1 + i] for i in range(5)], seq_len=m[6], sa_intv=m[7])
    return bwt, sa, meta


def refseq() -> np.ndarray:
    g = lib().mcref_genome_size()
    p = lib().mcref_refseq()
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(2 * g,)).copy()


def bwt_search(codes: np.ndarray, start: int, stop: int):
    ln, fr = C.c_int(), C.c_int()
    loc = np.zeros(64, dtype=np.uint64)
    lib().mcref_bwt_search(codes.ctypes.data, start, stop, C.byref(ln), C.byref(fr), loc.ctypes.data)
    return ln.value, fr.value, loc[:fr.value].copy()


def align(use_nw: bool, s1: bytes, s2: bytes):
    o1 = C.create_string_buffer(len(s1) + len(s2) + 2)
    o2 = C.create_string_buffer(len(s1) + len(s2) + 2)
    n = lib().mcref_align(int(use_nw), len(s1), s1, len(s2), s2, o1, o2)
    assert n >= 0, "reference produced strings of unequal length"
    return o1.value, o2.value


def parse_reads(blob: bytes, n_reads: int, paired: bool):
    """Inverse of serialise_read() in oracle/ref_shim.cpp -> list of dicts, est per chunk."""
    reads, p = [], 0
    for _ in range(n_reads):
        rlen, score, sub, best, nc = struct.unpack_from("<5i", blob, p); p += 20
        cands = []
        for _c in range(nc):
            cs, ori, pidx, nf = struct.unpack_from("<4i", blob, p); p += 16
            frags = []
            for _f in range(nf):
                simple, rpos = struct.unpack_from("<2i", blob, p); p += 8
                (gpos,) = struct.unpack_from("<q", blob, p); p += 8
                rl, gl, al = struct.unpack_from("<3i", blob, p); p += 12
                a1 = blob[p:p + al]; p += al
                a2 = blob[p:p + al]; p += al
                frags.append((simple, rpos, gpos, rl, gl, a1, a2))
            cands.append(dict(score=cs, orientation=ori, paired=pidx, frags=frags))
        reads.append(dict(rlen=rlen, score=score, sub_score=sub, best=best, cands=cands))
    est = []
    while p < len(blob):
        (e,) = struct.unpack_from("<i", blob, p); p += 4
        est.append(e)
    return reads, est


def map_reads(seq: np.ndarray, off: np.ndarray, paired: bool, update_profile: bool = True):
    n = C.c_int64()
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    p = lib().mcref_map(len(off) - 1, seq.ctypes.data, off.ctypes.data, int(paired), int(update_profile), C.byref(n))
    blob = _take(p, n)
    return parse_reads(blob, len(off) - 1, paired)


def set_avg_dist(v: int) -> None:
    """The value the next chunk derives its EstiDistance from (operator-level test of AlignmentRescue)."""
    lib().mcref_set_avg_dist(C.c_uint32(int(v)))


def counters():
    a = (C.c_int64 * 8)()
    lib().mcref_counters(a)
    return dict(zip(["reads", "mapped", "paired", "dist_sum", "len_sum", "avgDist", "avgReadLength", "FragmentSize"], list(a)))


def profile(beg: int = 0, end: int | None = None) -> np.ndarray:
    """int32 [n, 10]: A C G T multi_hit readCount F1 R2 F2 R1."""
    if end is None:
        end = lib().mcref_genome_size()
    out = np.zeros((end - beg, 10), dtype=np.int32)
    lib().mcref_profile(beg, end, out.ctypes.data)
    return out


def indels(which: int):
    n = C.c_int64()
    blob = _take(lib().mcref_indels(which, C.byref(n)), n)
    out, p = [], 0
    while p < len(blob):
        pos, cnt, ln = struct.unpack_from("<qii", blob, p); p += 16
        out.append((pos, blob[p:p + ln], cnt)); p += ln
    return out


def breakpoints():
    n = C.c_int64()
    blob = _take(lib().mcref_breakpoints(C.byref(n)), n)
    a = np.frombuffer(blob, dtype=np.int64).reshape(-1, 2)
    return [(int(x), int(y)) for x, y in a]


def sites(which: int):
    n = C.c_int64()
    blob = _take(lib().mcref_sites(which, C.byref(n)), n)
    a = np.frombuffer(blob, dtype=np.int64).reshape(-1, 2)
    return [(int(x), int(y)) for x, y in a]


# fields of Variant_t the reference defines per VarType (the others are stale locals of IdentifyVariants)
VARIANT_FIELDS = {0: ("DP", "AD_ref", "AD_alt", "GenoType", "qscore", "alt"), 1: ("DP", "AD_ref", "AD_alt", "GenoType", "qscore", "alt"),
                  2: ("DP", "AD_ref", "AD_alt", "GenoType", "qscore", "alt"), 5: ("DP",), 6: ("DP",), 10: ("DP", "AD_alt", "qscore"),
                  11: ("DP", "AD_ref", "GenoType", "qscore")}


def variant_scan(min_allele_depth=5, frequency_thr=0.2, somatic=0, gvcf=0, monomorphic=0, ploidy=2, min_cnv_size=50,
                 min_unmapped_size=50):
    """CalBlockReadDepth + IdentifyVariants (+ RemoveConsecutiveGenomicVariant) of the unmodified reference on the
    profile its mapping left behind -> (list of dicts like api.Context.variant_scan, BlockDepthArr int32)."""
    n = C.c_int64()
    p = lib().mcref_variant_scan(min_allele_depth, frequency_thr, somatic, gvcf, monomorphic, ploidy, min_cnv_size,
                                 min_unmapped_size, C.byref(n))
    b = _take(p, n)
    (nv,) = struct.unpack_from("<q", b, 0); o = 8
    out = []
    for _ in range(nv):
        g, t, dp, ar, aa, gt, q, al = struct.unpack_from("<q7i", b, o); o += 36
        out.append(dict(gPos=g, VarType=t, DP=dp, AD_ref=ar, AD_alt=aa, GenoType=gt, qscore=q, alt=b[o:o + al])); o += al
    (nb,) = struct.unpack_from("<i", b, o); o += 4
    return out, np.frombuffer(b, dtype="<i4", count=nb, offset=o).copy()


def variants_equal(mine, ref, limit=5):
    """Compares two variant lists on (gPos, VarType) and the fields the reference defines for the type."""
    bad = 0
    if len(mine) != len(ref):
        print("variant count", len(mine), len(ref)); bad += 1
    for a, r in zip(mine, ref):
        keys = ("gPos", "VarType") + VARIANT_FIELDS.get(r["VarType"], ())
        av = tuple(a[k].rstrip(b"\0") if k == "alt" else a[k] for k in keys)
        rv = tuple(r[k] for k in keys)
        if av != rv:
            bad += 1
            if bad <= limit:
                print("variant differs", dict(zip(keys, av)), dict(zip(keys, rv)))
    return bad == 0


def seed_cluster(seq: bytes):
    """IdentifySimplePairs + SimplePairClustering of the reference for one read (mcref_seed_cluster): (pairs, clusters) with
    pairs = [(rPos, gPos, len)] in the reference's sorted order and clusters = [(score, [(rPos, gPos, len)])]."""
    import struct
    n = C.c_int64()
    p = lib().mcref_seed_cluster(seq, len(seq), C.byref(n))
    b = _take(p, n)
    o = 0
    ns = struct.unpack_from("<i", b, o)[0]; o += 4
    sp = []
    for _ in range(ns):
        r, g, l = struct.unpack_from("<iqi", b, o); o += 16; sp.append((r, g, l))
    nc = struct.unpack_from("<i", b, o)[0]; o += 4
    cv = []
    for _ in range(nc):
        sc, nf = struct.unpack_from("<ii", b, o); o += 8
        fr = []
        for _ in range(nf):
            r, g, l = struct.unpack_from("<iqi", b, o); o += 16; fr.append((r, g, l))
        cv.append((sc, fr))
    return sp, cv
