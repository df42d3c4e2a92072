"""ctypes access to oracle/libmcoracle.so, the CPU restatement of the reference hot path
(oracle/restate/mapcaller_oracle.cpp).  Test infrastructure only: the checker of the CUDA path."""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "libmcoracle.so")

import ref_oracle  # shares the record parser: both libraries serialise reads identically


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.mco_create.restype = C.c_void_p
        L.mco_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.mco_destroy.argtypes = [C.c_void_p]
        L.mco_genome_size.restype = C.c_int64
        L.mco_genome_size.argtypes = [C.c_void_p]
        L.mco_map.restype = C.c_void_p
        L.mco_map.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.mco_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.mco_work.argtypes = [C.POINTER(C.c_int64)]
        L.mco_profile.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        for f in ("mco_indels", "mco_sites"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.mco_breakpoints.restype = C.c_void_p
        L.mco_breakpoints.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.mco_bwt_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        L.mco_align.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]
        L.mco_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _take(ptr, n) -> bytes:
    b = C.string_at(ptr, n.value)
    lib().mco_free(ptr)
    return b


class Oracle:
    def __init__(self, prefix: str, max_pos_diff=30, max_clip=5, max_dup=5, max_mismatch_rate=0.05, alg_ksw2=0, **_):
        self.h = lib().mco_create(prefix.encode(), max_pos_diff, max_clip, max_dup, max_mismatch_rate, int(not alg_ksw2))
        if not self.h:
            raise RuntimeError("oracle could not load index %s" % prefix)

    def close(self):
        if self.h:
            lib().mco_destroy(self.h)
            self.h = None

    def map_reads(self, seq, off, paired: bool, update_profile: bool = True):
        seq = np.ascontiguousarray(seq, dtype=np.uint8); off = np.ascontiguousarray(off, dtype=np.int64)
        n = C.c_int64()
        p = lib().mco_map(self.h, len(off) - 1, seq.ctypes.data, off.ctypes.data, int(paired), int(update_profile), C.byref(n))
        return ref_oracle.parse_reads(_take(p, n), len(off) - 1, paired)

    def counters(self):
        a = (C.c_int64 * 8)(); lib().mco_counters(self.h, a)
        return dict(zip(["reads", "mapped", "paired", "dist_sum", "len_sum", "avgDist"], list(a)[:6]))

    def work(self):
        a = (C.c_int64 * 5)(); lib().mco_work(a)
        return dict(zip(["seed_blocks", "locate_blocks", "sa_reads", "dp_cells", "dp_tasks"], list(a)))

    def profile(self, beg=0, end=None):
        if end is None:
            end = lib().mco_genome_size(self.h)
        out = np.zeros((end - beg, 10), dtype=np.int32)
        lib().mco_profile(self.h, beg, end, out.ctypes.data)
        return out

    def indels(self, which):
        n = C.c_int64(); blob = _take(lib().mco_indels(self.h, which, C.byref(n)), n)
        out, p = [], 0
        while p < len(blob):
            pos, cnt, ln = struct.unpack_from("<qii", blob, p); p += 16
            out.append((pos, blob[p:p + ln], cnt)); p += ln
        return out

    def breakpoints(self):
        n = C.c_int64(); a = np.frombuffer(_take(lib().mco_breakpoints(self.h, C.byref(n)), n), dtype=np.int64).reshape(-1, 2)
        return [(int(x), int(y)) for x, y in a]

    def sites(self, which):
        n = C.c_int64(); a = np.frombuffer(_take(lib().mco_sites(self.h, which, C.byref(n)), n), dtype=np.int64).reshape(-1, 2)
        return [(int(x), int(y)) for x, y in a]

    def bwt_search(self, codes: np.ndarray, start: int, stop: int):
        ln, fr = C.c_int(), C.c_int(); loc = np.zeros(64, dtype=np.uint64)
        lib().mco_bwt_search(self.h, codes.ctypes.data, start, stop, C.byref(ln), C.byref(fr), loc.ctypes.data)
        return ln.value, fr.value, loc[:fr.value].copy()


def align(use_nw: bool, s1: bytes, s2: bytes):
    o1 = C.create_string_buffer(len(s1) + len(s2) + 2); o2 = C.create_string_buffer(len(s1) + len(s2) + 2)
    n = lib().mco_align(int(use_nw), len(s1), s1, len(s2), s2, o1, o2)
    assert n >= 0
    return o1.value, o2.value
