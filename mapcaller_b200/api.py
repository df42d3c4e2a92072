"""ctypes binding of libmapcaller_b200.so (include/mapcaller_b200.h).

This is plumbing for the tests and bench.py: the product is the C-ABI library.  There is no Python or
CPU implementation behind these classes; if the shared library (built by __graft_entry__.build()) is
missing, importing fails loudly, and every mapping call fails with MC_ERR_CUDA on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmapcaller_b200.so")

CHUNK_READS = 200


class McError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("paired", C.c_int32), ("alg_ksw2", C.c_int32), ("max_pos_diff", C.c_int32), ("max_clip", C.c_int32),
                ("max_dup", C.c_int32), ("max_mismatch_rate", C.c_float), ("update_profile", C.c_int32),
                ("want_alignments", C.c_int32), ("device", C.c_int32), ("shard_rank", C.c_int32),
                ("shard_count", C.c_int32), ("reserved", C.c_int32 * 5)]


class IndexView(C.Structure):
    _fields_ = [("bwt", C.c_void_p), ("bwt_size", C.c_uint64), ("primary", C.c_uint64), ("L2", C.c_uint64 * 5),
                ("seq_len", C.c_uint64), ("sa", C.c_void_p), ("n_sa", C.c_uint64), ("sa_intv", C.c_int32),
                ("pac", C.c_void_p), ("genome_size", C.c_int64), ("n_chrom", C.c_int32), ("chrom_len", C.c_void_p),
                ("chrom_name", C.c_void_p)]


class BatchIn(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("seq", C.c_void_p), ("seq_off", C.c_void_p)]


class BatchOut(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_cands", C.c_int64), ("n_frags", C.c_int64), ("n_aln_bytes", C.c_int64),
                ("n_pairs", C.c_int64), ("n_chunks", C.c_int64), ("reads", C.c_void_p), ("cands", C.c_void_p),
                ("frags", C.c_void_p), ("aln", C.c_void_p), ("pairs", C.c_void_p),
                ("chunks", C.c_void_p), ("replays", C.c_int32)]


class Totals(C.Structure):
    _fields_ = [("total_reads", C.c_int64), ("total_mapped", C.c_int64), ("total_paired", C.c_int64),
                ("total_distance", C.c_int64), ("read_length_sum", C.c_int64), ("avg_dist", C.c_uint32), ("pad", C.c_uint32)]


class VcParams(C.Structure):
    """mc_vc_params (include/mapcaller_b200.h): thresholds of the variant-calling scan."""
    _fields_ = [("min_allele_depth", C.c_int32), ("frequency_thr", C.c_float)] + \
               [(k, C.c_int32) for k in ("somatic", "gvcf", "monomorphic", "ploidy", "min_cnv_size", "min_unmapped_size")]


VARIANT_DT = np.dtype([("gPos", "<i8"), ("rec0", "<u8"), ("rec1", "<u8"), ("alt_off", "<i4"), ("alt_len", "<i4"),
                       ("DP", "<u2"), ("AD_ref", "<u2"), ("AD_alt", "<u2"), ("GenoType", "u1"), ("qscore", "u1"),
                       ("VarType", "u1"), ("alt", "S3"), ("pad", "V4")])
assert VARIANT_DT.itemsize == 48


SAM_DT = np.dtype([("pos", "<i8"), ("mate_pos", "<i8")] + [(k, "<i4") for k in ("flag", "chrom", "mapq", "tlen", "nm", "as", "xs", "cigar_off",
                                                                                 "cigar_len", "reverse", "has_mate", "pad")])
assert SAM_DT.itemsize == 64
_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def complement_read(seq: bytes) -> bytes:
    """GetComplementarySeq of the reference (src/tools.cpp:5-29): reverse complement, anything but ACGT/acgt -> N."""
    return bytes(_COMP[b] if chr(b) in "ACGTacgt" else ord("N") for b in reversed(seq))


def format_sam_line(rec, cigar: bytes, name: bytes, seq: bytes, qual: bytes | None, chrom_names) -> bytes | None:
    """One SAM line from an mc_sam_rec, as GeneratePairedSamStream / GenerateSingleSamStream print it (src/SamReport.cpp:332,346-363,401-434)."""
    if rec["flag"] < 0:
        return None
    if rec["reverse"]:
        seq, qual = complement_read(seq), (qual[::-1] if qual is not None else None)
    q = qual if qual is not None else b"*"
    if rec["chrom"] < 0:
        return b"%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\tAS:i:0\tXS:i:0" % (name, rec["flag"], seq, q)
    mate = b"=\t%d\t%d" % (rec["mate_pos"], rec["tlen"]) if rec["has_mate"] else b"*\t0\t0"
    return b"%s\t%d\t%s\t%d\t%d\t%s\t%s\t%s\t%s\tNM:i:%d\tAS:i:%d\tXS:i:%d" % (
        name, rec["flag"], chrom_names[rec["chrom"]], rec["pos"], rec["mapq"], cigar, mate, seq, q, rec["nm"], rec["as"], rec["xs"])


class Stats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("ms_seed", "ms_locate", "ms_cluster", "ms_pair", "ms_align", "ms_profile",
                                           "ms_h2d", "ms_d2h", "ms_total")] + \
               [(k, C.c_int64) for k in ("seed_blocks", "locate_blocks", "sa_reads", "dp_cells", "dp_tasks",
                                          "profile_columns", "profile_atomics", "kernel_launches")] + [("ms_reduce", C.c_double)] + \
               [("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("ms_reset", C.c_double)]


READ_DT = np.dtype([("score", "<i4"), ("sub_score", "<i4"), ("best_idx", "<i4"), ("cand_begin", "<i4"), ("n_cand", "<i4"), ("rlen", "<i4")])
CAND_DT = np.dtype([("score", "<i4"), ("orientation", "<i4"), ("paired_idx", "<i4"), ("frag_begin", "<i4"), ("n_frag", "<i4"), ("pad", "<i4")])
FRAG_DT = np.dtype([("gPos", "<i8"), ("rPos", "<i4"), ("rLen", "<i4"), ("gLen", "<i4"), ("bSimple", "<i4"), ("aln_off", "<i4"),
                    ("aln_len", "<i4"), ("aln_cap", "<i4"), ("pad", "<i4")])
PAIR_DT = np.dtype([("gPos1", "<i8"), ("gPos2", "<i8"), ("dist", "<i8")])
CHUNK_DT = np.dtype([("n_reads", "<i4"), ("mapped", "<i4"), ("paired", "<i4"), ("est_distance", "<i4"), ("dist_sum", "<i8"), ("len_sum", "<i8")])
INDEL_DT = np.dtype([("pos", "<i8"), ("kind", "<i4"), ("len", "<i4"), ("count", "<i4"), ("seq_off", "<i4")])
PROFILE_DT = np.dtype([("bits", "<u8"), ("F1", "<u2"), ("R2", "<u2"), ("F2", "<u2"), ("R1", "<u2")])

_lib = None


def lib():
    """Loads the C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise McError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first" % _LIB_PATH)
        L = C.CDLL(_LIB_PATH)
        L.mc_last_error.restype = C.c_char_p
        L.mc_version.restype = C.c_char_p
        L.mc_index_build.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.mc_index_build_gpu.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.mc_index_build_fasta.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.mc_index_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.mc_index_save.argtypes = [C.c_void_p, C.c_char_p]
        L.mc_index_wrap.argtypes = [C.POINTER(IndexView), C.POINTER(C.c_void_p)]
        L.mc_index_get.argtypes = [C.c_void_p, C.POINTER(IndexView)]
        L.mc_index_free.argtypes = [C.c_void_p]
        L.mc_params_default.argtypes = [C.POINTER(Params)]
        L.mc_ctx_create.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(C.c_void_p)]
        L.mc_ctx_destroy.argtypes = [C.c_void_p]
        L.mc_map_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(BatchOut)]
        L.mc_stage_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.c_int32]
        L.mc_stage_batch_async.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.c_int32]
        L.mc_ingest_fastq.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.mc_map_staged.argtypes = [C.c_void_p, C.c_int32, C.POINTER(BatchOut)]
        L.mc_ingest_prefetch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.mc_sam_text.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.mc_get_totals.argtypes = [C.c_void_p, C.POINTER(Totals)]
        L.mc_set_totals.argtypes = [C.c_void_p, C.POINTER(Totals)]
        L.mc_reset.argtypes = [C.c_void_p]
        L.mc_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.mc_host_free.argtypes = [C.c_void_p]
        L.mc_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.mc_reset_stats.argtypes = [C.c_void_p]
        L.mc_profile_read.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.mc_profile_summary.argtypes = [C.c_void_p, C.c_void_p]
        L.mc_profile_checksum.argtypes = [C.c_void_p, C.c_void_p]
        L.mc_begin_library.argtypes = [C.c_void_p]
        L.mc_profile_indels.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p)]
        L.mc_profile_breakpoints.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.mc_profile_sites.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.mc_vc_params_default.argtypes = [C.POINTER(VcParams)]
        L.mc_vc_params_default.restype = None
        L.mc_variant_scan.argtypes = [C.c_void_p, C.POINTER(VcParams), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.mc_sam_records.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p)]
        L.mc_profile_allreduce.argtypes = [C.c_void_p, C.c_void_p]
        L.mc_profile_reduce_scatter.argtypes = [C.c_void_p]
        L.mc_read_alignment_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mc_defer_profile.argtypes = [C.c_void_p, C.c_int32]
        L.mc_update_profile_last.argtypes = [C.c_void_p]
        L.mc_rescue_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.mc_profile_owned.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.mc_comm_unique_id.argtypes = [C.c_void_p]
        L.mc_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.mc_align_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int64] + [C.c_void_p] * 8
        L.mc_bwt_search_batch.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
        L.mc_seed_cluster_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.c_void_p]
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise McError("%s failed (%d): %s" % (what, rc, lib().mc_last_error().decode()))


def _view(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * (count * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


def pinned_array(shape, dtype) -> np.ndarray:
    """numpy array backed by page-locked memory from mc_host_alloc (kept alive for the life of the process)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _check(lib().mc_host_alloc(max(n, 1), C.byref(p)), "mc_host_alloc")
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Index:
    """mc_index: the reference's bwaidx_t image (reference src/structure.h:32-72)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def build(cls, fwd_codes: np.ndarray, chrom_len=None, chrom_name=None, threads: int = 0, gpu_device=None) -> "Index":
        fwd_codes = np.ascontiguousarray(fwd_codes, dtype=np.uint8)
        h = C.c_void_p()
        if chrom_len is None:
            chrom_len = [len(fwd_codes)]
        lens = np.asarray(chrom_len, dtype=np.int32)
        names = None
        if chrom_name is not None:
            names = (C.c_char_p * len(chrom_name))(*[n.encode() for n in chrom_name])
        if gpu_device is not None:   # suffix array sorted on the GPU (same index, same files)
            _check(lib().mc_index_build_gpu(fwd_codes.ctypes.data, len(fwd_codes), len(lens), lens.ctypes.data, names, int(gpu_device), C.byref(h)), "mc_index_build_gpu")
            return cls(h)
        _check(lib().mc_index_build(fwd_codes.ctypes.data, len(fwd_codes), len(lens), lens.ctypes.data, names, threads, C.byref(h)), "mc_index_build")
        return cls(h)

    @classmethod
    def build_fasta(cls, path: str, threads: int = 0) -> "Index":
        h = C.c_void_p()
        _check(lib().mc_index_build_fasta(path.encode(), threads, C.byref(h)), "mc_index_build_fasta")
        return cls(h)

    @classmethod
    def load(cls, prefix: str) -> "Index":
        h = C.c_void_p()
        _check(lib().mc_index_load(prefix.encode(), C.byref(h)), "mc_index_load")
        return cls(h)

    def save(self, prefix: str) -> None:
        _check(lib().mc_index_save(self._h, prefix.encode()), "mc_index_save")

    def view(self) -> IndexView:
        v = IndexView()
        _check(lib().mc_index_get(self._h, C.byref(v)), "mc_index_get")
        return v

    def view_arrays(self) -> dict:
        """The arrays of the image as numpy views (valid while the index lives): bwt words, sampled SA, pac bytes."""
        v = self.view()
        return dict(primary=int(v.primary), L2=[int(x) for x in v.L2], seq_len=int(v.seq_len),
                    bwt=_view(v.bwt, int(v.bwt_size), np.dtype("<u4")), sa=_view(v.sa, int(v.n_sa), np.dtype("<u8")),
                    pac=_view(v.pac, int(v.genome_size // 4 + 1), np.dtype("u1")))

    @property
    def genome_size(self) -> int:
        return self.view().genome_size

    def close(self):
        if self._h:
            lib().mc_index_free(self._h)
            self._h = None


class Context:
    """mc_ctx: one GPU's mapping state; map_batch() replaces the body of the reference's ReadMapping()."""

    def __init__(self, index: Index, **kw):
        p = Params()
        lib().mc_params_default(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError("unknown parameter %s" % k)
            setattr(p, k, v)
        self.params = p
        self._index = index
        h = C.c_void_p()
        _check(lib().mc_ctx_create(index._h, C.byref(p), C.byref(h)), "mc_ctx_create")
        self._h = h

    def close(self):
        if self._h:
            lib().mc_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def _batch(seq: np.ndarray, off: np.ndarray):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        b = BatchIn(len(off) - 1, seq.ctypes.data, off.ctypes.data)
        return b, (seq, off)

    def _wrap(self, out: BatchOut, copy: bool):
        f = (lambda x: x.copy()) if copy else (lambda x: x)
        return dict(reads=f(_view(out.reads, out.n_reads if out.reads else 0, READ_DT)), cands=f(_view(out.cands, out.n_cands, CAND_DT)),
                    frags=f(_view(out.frags, out.n_frags, FRAG_DT)), aln=f(_view(out.aln, out.n_aln_bytes, np.dtype("u1"))),
                    pairs=f(_view(out.pairs, out.n_pairs, PAIR_DT)), chunks=f(_view(out.chunks, out.n_chunks, CHUNK_DT)),
                    replays=out.replays)

    def map_batch(self, seq: np.ndarray, off: np.ndarray, copy: bool = True):
        b, keep = self._batch(seq, off)
        out = BatchOut()
        _check(lib().mc_map_batch(self._h, C.byref(b), C.byref(out)), "mc_map_batch")
        return self._wrap(out, copy)

    def read_alignment_batch(self, seq: np.ndarray, off: np.ndarray):
        """mc_read_alignment_batch: SimplePairClustering .. ProduceReadAlignment for independent reads (the context's state is untouched)."""
        b, keep = self._batch(seq, off)
        out = BatchOut()
        _check(lib().mc_read_alignment_batch(self._h, C.byref(b), C.byref(out)), "mc_read_alignment_batch")
        return self._wrap(out, True)

    def rescue_batch(self, seq: np.ndarray, off: np.ndarray, avg_dist: int):
        """mc_rescue_batch: the pairs of the batch with EstiDistance = (int)(avg_dist * 1.5) for every pair, no feedback."""
        b, keep = self._batch(seq, off)
        out = BatchOut()
        _check(lib().mc_rescue_batch(self._h, C.byref(b), C.c_uint32(int(avg_dist)), C.byref(out)), "mc_rescue_batch")
        return self._wrap(out, True)

    def defer_profile(self, on: bool = True):
        _check(lib().mc_defer_profile(self._h, int(on)), "mc_defer_profile")

    def update_profile_last(self):
        """UpdateProfile / UpdateMultiHitCount for the reads of the batch just mapped (after defer_profile())."""
        _check(lib().mc_update_profile_last(self._h), "mc_update_profile_last")

    def stage_batch(self, seq: np.ndarray, off: np.ndarray, slot: int = 0):
        b, keep = self._batch(seq, off)
        _check(lib().mc_stage_batch(self._h, C.byref(b), slot), "mc_stage_batch")

    def stage_batch_async(self, seq: np.ndarray, off: np.ndarray, slot: int = 0):
        """Queues the copy of a batch into device slot `slot` and returns; the arrays must stay untouched (and alive: they are
        kept referenced here) until the next map_staged(slot) has returned."""
        b, keep = self._batch(seq, off)
        self._staging = getattr(self, "_staging", {}); self._staging[slot] = keep
        _check(lib().mc_stage_batch_async(self._h, C.byref(b), slot), "mc_stage_batch_async")

    def ingest_fastq(self, text1, text2=None, slot: int = 0, max_reads: int = 0, final: bool = True, keep_text: bool = False, fasta: bool = False) -> dict:
        """Parses blocks of FASTQ text (bytes / uint8 arrays; one per mate file, or one with adjacent mates) on the device into
        slot `slot`; returns n_reads, consumed1, consumed2, n_bases.  map_staged(slot) maps the batch.  fasta: the text is FASTA
        with one line of bases per record."""
        t1 = np.frombuffer(text1, dtype=np.uint8) if isinstance(text1, (bytes, bytearray)) else np.ascontiguousarray(text1, dtype=np.uint8)
        t2 = None if text2 is None else (np.frombuffer(text2, dtype=np.uint8) if isinstance(text2, (bytes, bytearray)) else np.ascontiguousarray(text2, dtype=np.uint8))
        arg = (C.c_int64 * 6)(t1.ctypes.data, len(t1), t2.ctypes.data if t2 is not None else 0, len(t2) if t2 is not None else 0, max_reads, int(bool(final)) | (int(bool(fasta)) << 32))   # final_block, format
        out = (C.c_int64 * 6)()
        _check(lib().mc_ingest_fastq(self._h, arg, slot, out), "mc_ingest_fastq")
        return dict(n_reads=int(out[0]), consumed1=int(out[1]), consumed2=int(out[2]), n_bases=int(out[3]), records1=int(out[4]), records2=int(out[5]))

    @staticmethod
    def _fastq_arg(text1, text2, max_reads, final, fasta=False):
        t1 = np.frombuffer(text1, dtype=np.uint8) if isinstance(text1, (bytes, bytearray)) else np.ascontiguousarray(text1, dtype=np.uint8)
        t2 = None if text2 is None else (np.frombuffer(text2, dtype=np.uint8) if isinstance(text2, (bytes, bytearray)) else np.ascontiguousarray(text2, dtype=np.uint8))
        arg = (C.c_int64 * 6)(t1.ctypes.data, len(t1), t2.ctypes.data if t2 is not None else 0, len(t2) if t2 is not None else 0, max_reads, int(bool(final)) | (int(bool(fasta)) << 32))   # final_block, format
        return arg, (t1, t2)

    def ingest_prefetch(self, text1, text2=None, slot: int = 0):
        """Queues the host -> device copy of the blocks the next ingest_fastq(slot) parses (page-locked arrays) and returns."""
        arg, keep = self._fastq_arg(text1, text2, 0, True)
        self._prefetched = getattr(self, "_prefetched", {}); self._prefetched[slot] = keep
        _check(lib().mc_ingest_prefetch(self._h, arg, slot), "mc_ingest_prefetch")

    def sam_text_raw(self, slot: int, all_best: bool = False) -> int:
        """mc_sam_text: the SAM lines of the batch just mapped from `slot` are left in the library's page-locked buffer; returns
        the number of bytes."""
        p, n = C.c_void_p(), C.c_int64()
        _check(lib().mc_sam_text(self._h, slot, int(all_best), C.byref(p), C.byref(n)), "mc_sam_text")
        self._sam_text = (p.value, n.value)
        return int(n.value)

    def sam_text(self, slot: int, all_best: bool = False) -> bytes:
        n = self.sam_text_raw(slot, all_best)
        return C.string_at(self._sam_text[0], n)

    def map_staged(self, slot: int = 0, copy: bool = False):
        out = BatchOut()
        _check(lib().mc_map_staged(self._h, slot, C.byref(out)), "mc_map_staged")
        return self._wrap(out, copy)

    def align_batch(self, pairs, ksw2: bool = False):
        """pairs: list of (read piece bytes, genome piece bytes) -> list of (aln1, aln2) through the DP kernel."""
        n = len(pairs)
        s1 = np.frombuffer(b"".join(p[0] for p in pairs), dtype=np.uint8)
        s2 = np.frombuffer(b"".join(p[1] for p in pairs), dtype=np.uint8)
        o1 = np.zeros(n + 1, dtype=np.int64); o1[1:] = np.cumsum([len(p[0]) for p in pairs])
        o2 = np.zeros(n + 1, dtype=np.int64); o2[1:] = np.cumsum([len(p[1]) for p in pairs])
        oo = o1 + o2
        out1 = np.zeros(int(oo[-1]) + 1, dtype=np.uint8); out2 = np.zeros(int(oo[-1]) + 1, dtype=np.uint8)
        ln = np.zeros(n, dtype=np.int32)
        _check(lib().mc_align_batch(self._h, int(ksw2), n, s1.ctypes.data, o1.ctypes.data, s2.ctypes.data, o2.ctypes.data, oo.ctypes.data,
                                    out1.ctypes.data, out2.ctypes.data, ln.ctypes.data), "mc_align_batch")
        return [(out1[oo[i]:oo[i] + ln[i]].tobytes(), out2[oo[i]:oo[i] + ln[i]].tobytes()) for i in range(n)]

    def bwt_search_batch(self, codes: np.ndarray, off: np.ndarray, start: np.ndarray):
        """BWT_Search for n (codes[off[i]:off[i+1]], start[i]) queries -> (len[n], freq[n], list of sorted location arrays)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8); off = np.ascontiguousarray(off, dtype=np.int64)
        start = np.ascontiguousarray(start, dtype=np.int32)
        n = len(start)
        ln = np.zeros(n, dtype=np.int32); fr = np.zeros(n, dtype=np.int32); loc = np.zeros((n, 50), dtype=np.uint64)
        _check(lib().mc_bwt_search_batch(self._h, n, codes.ctypes.data, off.ctypes.data, start.ctypes.data, ln.ctypes.data, fr.ctypes.data,
                                         loc.ctypes.data), "mc_bwt_search_batch")
        return ln, fr, [np.sort(loc[i, :fr[i]]) for i in range(n)]

    def seed_cluster_batch(self, seq: np.ndarray, off: np.ndarray):
        """IdentifySimplePairs + SimplePairClustering per read (taken as given): list of (pairs, clusters) with
        pairs = [(rPos, gPos, len)] sorted by (PosDiff, rPos) and clusters = [(score, [(rPos, gPos, len)])]."""
        b, keep = self._batch(seq, off)
        out = (C.c_int64 * 7)()
        _check(lib().mc_seed_cluster_batch(self._h, C.byref(b), out), "mc_seed_cluster_batch")
        n = int(out[0])
        if n == 0:
            return []
        poff = _view(out[1], n + 1, np.dtype("<i8")); npair = _view(out[2], n, np.dtype("<i4")); nclu = _view(out[5], n, np.dtype("<i4"))
        total = int(poff[n])
        pairs = _view(out[3], total, np.dtype([("gPos", "<i8"), ("rPos", "<i4"), ("len", "<i4")]))
        clus = _view(out[6], total, np.dtype([("score", "<i4"), ("b", "<i4"), ("e", "<i4")]))
        res = []
        for r in range(n):
            o = int(poff[r])
            ps = [(int(p["rPos"]), int(p["gPos"]), int(p["len"])) for p in pairs[o:o + int(npair[r])]]
            cs = [(int(c["score"]), [(int(p["rPos"]), int(p["gPos"]), int(p["len"])) for p in pairs[int(c["b"]):int(c["e"])]]) for c in clus[o:o + int(nclu[r])]]
            res.append((ps, cs))
        return res

    def totals(self) -> dict:
        t = Totals()
        _check(lib().mc_get_totals(self._h, C.byref(t)), "mc_get_totals")
        return {k: getattr(t, k) for k, _ in Totals._fields_ if k != "pad"}

    def set_totals(self, **kw):
        t = Totals()
        _check(lib().mc_get_totals(self._h, C.byref(t)), "mc_get_totals")
        for k, v in kw.items():
            setattr(t, k, v)
        _check(lib().mc_set_totals(self._h, C.byref(t)), "mc_set_totals")

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().mc_comm_unique_id(buf), "mc_comm_unique_id")
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, n_ranks: int):
        _check(lib().mc_comm_init(self._h, uid, rank, n_ranks), "mc_comm_init")

    def profile_allreduce(self):
        _check(lib().mc_profile_allreduce(self._h, None), "mc_profile_allreduce")

    def profile_reduce_scatter(self):
        """Every rank keeps the library's counters of one genome tile; variant_scan / profile_summary / profile_checksum are
        collectives afterwards (include/mapcaller_b200.h)."""
        _check(lib().mc_profile_reduce_scatter(self._h), "mc_profile_reduce_scatter")

    def profile_owned(self):
        b, e = C.c_int64(), C.c_int64()
        _check(lib().mc_profile_owned(self._h, C.byref(b), C.byref(e)), "mc_profile_owned")
        return b.value, e.value

    def reset(self):
        _check(lib().mc_reset(self._h), "mc_reset")

    def begin_library(self):
        _check(lib().mc_begin_library(self._h), "mc_begin_library")

    def profile_checksum(self):
        v = (C.c_uint64 * 2)()
        _check(lib().mc_profile_checksum(self._h, v), "mc_profile_checksum")
        return int(v[0]), int(v[1])

    def stats(self) -> dict:
        s = Stats()
        _check(lib().mc_get_stats(self._h, C.byref(s)), "mc_get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def reset_stats(self):
        _check(lib().mc_reset_stats(self._h), "mc_reset_stats")

    def profile(self, beg: int = 0, end: int | None = None) -> np.ndarray:
        """MappingRecord_t records (reference src/structure.h:152-163) of [beg, end)."""
        if end is None:
            end = self._index.genome_size
        out = np.zeros(end - beg, dtype=PROFILE_DT)
        _check(lib().mc_profile_read(self._h, beg, end, out.ctypes.data), "mc_profile_read")
        return out

    def profile_columns(self, beg: int = 0, end: int | None = None) -> np.ndarray:
        """int32 [n, 10]: A C G T multi_hit readCount F1 R2 F2 R1 (unpacked MappingRecord_t)."""
        p = self.profile(beg, end)
        b = p["bits"]
        cols = [(b >> np.uint64(s)) & np.uint64(0xFFF) for s in (0, 12, 24, 36, 48)] + [(b >> np.uint64(60)) & np.uint64(0xF)]
        cols += [p["F1"], p["R2"], p["F2"], p["R1"]]
        return np.stack([c.astype(np.int32) for c in cols], axis=1)

    def profile_summary(self) -> dict:
        """CheckMappingCoverage / ReportDuplicationRate figures from device reductions (no profile download)."""
        v = (C.c_int64 * 4)()
        _check(lib().mc_profile_summary(self._h, v), "mc_profile_summary")
        return dict(aligned_bases=int(v[0]), coverage_sum=int(v[1]), dup_sites=int(v[2]), dup_reads=int(v[3]))

    def indels(self):
        """[(pos, seq bytes, count)] for insertions, same for deletions (InsertSeqMap / DeleteSeqMap)."""
        recs, n, arena = C.c_void_p(), C.c_int64(), C.c_void_p()
        _check(lib().mc_profile_indels(self._h, C.byref(recs), C.byref(n), C.byref(arena)), "mc_profile_indels")
        r = _view(recs.value, n.value, INDEL_DT)
        out = ([], [])
        for x in r:
            s = C.string_at(arena.value + int(x["seq_off"]), int(x["len"]))
            out[int(x["kind"])].append((int(x["pos"]), s, int(x["count"])))
        return out

    def sam_records(self):
        """SAM fields of the batch just mapped (mc_sam_records): (structured array SAM_DT, list of CIGAR byte strings)."""
        recs, n, arena = C.c_void_p(), C.c_int64(), C.c_void_p()
        _check(lib().mc_sam_records(self._h, C.byref(recs), C.byref(n), C.byref(arena)), "mc_sam_records")
        r = _view(recs.value, n.value, SAM_DT).copy()
        total = int((r["cigar_off"] + r["cigar_len"]).max()) if n.value else 0
        text = C.string_at(arena.value, total)
        return r, [text[o:o + l] if l else b"*" for o, l in zip(r["cigar_off"].tolist(), r["cigar_len"].tolist())]

    def variant_scan(self, **kw):
        """CalBlockReadDepth + IdentifyVariants of the reference (src/VariantCalling.cpp:106-120, 550-680) on the device.
        -> (list of dicts {gPos, VarType, DP, AD_ref, AD_alt, GenoType, qscore, alt, record}, block depths int32)."""
        vp = VcParams()
        lib().mc_vc_params_default(C.byref(vp))
        for k, v in kw.items():
            if not hasattr(vp, k):
                raise TypeError("unknown variant-scan parameter %r" % k)
            setattr(vp, k, v)
        recs, n, arena, depth, nb = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_void_p(), C.c_int64()
        _check(lib().mc_variant_scan(self._h, C.byref(vp), C.byref(recs), C.byref(n), C.byref(arena), C.byref(depth), C.byref(nb)),
               "mc_variant_scan")
        r = _view(recs.value, n.value, VARIANT_DT)
        out = []
        for x in r:
            t = int(x["VarType"])
            alt = C.string_at(arena.value + int(x["alt_off"]), int(x["alt_len"])) if t in (1, 2) else bytes(x["alt"])
            out.append(dict(gPos=int(x["gPos"]), VarType=t, DP=int(x["DP"]), AD_ref=int(x["AD_ref"]), AD_alt=int(x["AD_alt"]),
                            GenoType=int(x["GenoType"]), qscore=int(x["qscore"]), alt=alt, record=(int(x["rec0"]), int(x["rec1"]))))
        return out, _view(depth.value, nb.value, np.dtype("<i4")).copy()

    def variant_scan_raw(self, **kw):
        """mc_variant_scan without unpacking: the records stay in the library's host buffers; returns (n_records, n_blocks)."""
        vp = VcParams()
        lib().mc_vc_params_default(C.byref(vp))
        for k, v in kw.items():
            setattr(vp, k, v)
        recs, n, arena, depth, nb = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_void_p(), C.c_int64()
        _check(lib().mc_variant_scan(self._h, C.byref(vp), C.byref(recs), C.byref(n), C.byref(arena), C.byref(depth), C.byref(nb)),
               "mc_variant_scan")
        return int(n.value), int(nb.value)

    def breakpoints(self):
        recs, n = C.c_void_p(), C.c_int64()
        _check(lib().mc_profile_breakpoints(self._h, C.byref(recs), C.byref(n)), "mc_profile_breakpoints")
        a = _view(recs.value, n.value, np.dtype([("pos", "<i8"), ("count", "<i8")]))
        return [(int(x["pos"]), int(x["count"])) for x in a]

    def sites(self, kind: int):
        recs, n = C.c_void_p(), C.c_int64()
        _check(lib().mc_profile_sites(self._h, kind, C.byref(recs), C.byref(n)), "mc_profile_sites")
        a = _view(recs.value, n.value, np.dtype([("gPos", "<i8"), ("dist", "<i8")]))
        return [(int(x["gPos"]), int(x["dist"])) for x in a]


def unpack_reads(res: dict):
    """Batch result -> list of dicts shaped like tests/ref_oracle.parse_reads() output (for comparisons)."""
    reads, cands, frags, aln = res["reads"], res["cands"], res["frags"], res["aln"]
    out = []
    alnb = aln.tobytes()
    for r in reads:
        cl = []
        for c in cands[r["cand_begin"]:r["cand_begin"] + r["n_cand"]]:
            fl = []
            if c["score"] > 0:
                for f in frags[c["frag_begin"]:c["frag_begin"] + c["n_frag"]]:
                    if f["bSimple"]:
                        a1 = a2 = b""
                    else:
                        o, ln, cp = int(f["aln_off"]), int(f["aln_len"]), int(f["aln_cap"])
                        a1, a2 = alnb[o:o + ln], alnb[o + cp:o + cp + ln]
                    fl.append((int(f["bSimple"]), int(f["rPos"]), int(f["gPos"]), int(f["rLen"]), int(f["gLen"]), a1, a2))
            cl.append(dict(score=int(c["score"]), orientation=int(c["orientation"]) if c["score"] > 0 else -1, paired=int(c["paired_idx"]), frags=fl))
        out.append(dict(rlen=int(r["rlen"]), score=int(r["score"]), sub_score=int(r["sub_score"]), best=int(r["best_idx"]), cands=cl))
    return out
