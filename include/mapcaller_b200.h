/* mapcaller_b200.h -- C ABI of libmapcaller_b200.so
 *
 * B200-native (sm_100a) replacement for MapCaller's read-mapping hot path.  The reference has no
 * plugin/FFI layer: its de-facto operator interface is the set of `extern` C++ functions and
 * process-wide globals of reference src/structure.h:197-292.  Every entry
 * point below names the reference interface it replaces.  Plain pointers and sizes only: no STL,
 * no exceptions, no torch types.  All functions return 0 on success and a negative code on
 * failure; mc_last_error() then holds a message.  There is NO CPU fallback: every mc_ctx_* call
 * fails with MC_ERR_CUDA when no sm_100 device / driver is usable.
 *
 * Threading: one mc_ctx per GPU; calls on one ctx must be serialised by the caller (the reference
 * serialises the same state behind OutputLock/ProfileLock, src/ReadMapping.cpp:537,564).
 * Ownership: inputs are caller-owned and may be released when the call returns; output pointers
 * are library-owned pinned host memory, valid until the next mc_map_batch/mc_ctx_destroy on the ctx.
 */
#ifndef MAPCALLER_B200_H
#define MAPCALLER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_OK            0
#define MC_ERR_ARG      -1
#define MC_ERR_CUDA     -2
#define MC_ERR_IO       -3
#define MC_ERR_OVERFLOW -4
#define MC_ERR_NCCL     -5

#define MC_CHUNK_READS 200   /* ReadChunkSize, reference src/structure.h:24 */
#define MC_MAX_OCC      50   /* OCC_Thr, reference src/bwt_search.cpp:3 */
#define MC_MIN_SEED     16   /* MinSeedLength, reference src/structure.h:23 */
#define MC_MAX_RLEN   4000   /* longest read accepted (keeps the 16-bit profile counters exact) */

const char *mc_last_error(void);
const char *mc_version(void);

/* ------------------------------------------------------------------------------------------------
 * Index: the in-memory image of bwaidx_t / bwt_t (reference src/structure.h:32-72) as loaded by
 * bwa_idx_load + RestoreReferenceInfo (reference src/bwt_index.cpp:150,232).  On-disk format is the
 * reference's (.bwt .sa .pac .ann .amb; reference src/BWT_Index/bwtindex.c:53-149, bwt.c:174-196,
 * bntseq.c:59-91) so an index built by either side loads in the other.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mc_index mc_index;

typedef struct {
    const uint32_t *bwt;      /* interleaved occ+BWT words, 64-byte blocks (bwt_t::bwt) */
    uint64_t bwt_size;        /* number of uint32 words (bwt_t::bwt_size) */
    uint64_t primary;         /* bwt_t::primary */
    uint64_t L2[5];           /* bwt_t::L2 */
    uint64_t seq_len;         /* bwt_t::seq_len == 2 * genome_size */
    const uint64_t *sa;       /* sampled suffix array, sa[0] == (uint64_t)-1 (bwt_t::sa) */
    uint64_t n_sa;            /* bwt_t::n_sa */
    int32_t sa_intv;          /* bwt_t::sa_intv (32) */
    const uint8_t *pac;       /* forward-only 2-bit packed reference (bwaidx_t::pac) */
    int64_t genome_size;      /* GenomeSize == bntseq_t::l_pac */
    int32_t n_chrom;          /* iChromsomeNum */
    const int32_t *chrom_len; /* ChromosomeVec[i].len */
    const char *const *chrom_name; /* ChromosomeVec[i].name (may be NULL) */
} mc_index_view;

/* Replaces bwa_idx_build() (reference src/BWT_Index/bwtindex.c:77): builds the FM-index of
 * fwd + revcomp from 2-bit codes (0..3) of the concatenated contigs. */
int mc_index_build(const uint8_t *fwd_codes, int64_t genome_size, int32_t n_chrom, const int32_t *chrom_len,
                   const char *const *chrom_name, int32_t n_threads, mc_index **out);
/* The same with the suffix array sorted on a GPU (prefix doubling with radix sorts, 36 bytes of HBM per text symbol; texts
 * below 2^32 symbols): identical index, identical files. */
int mc_index_build_gpu(const uint8_t *fwd_codes, int64_t genome_size, int32_t n_chrom, const int32_t *chrom_len,
                       const char *const *chrom_name, int32_t device, mc_index **out);
/* FASTA front end of the same: N / IUPAC bases are replaced exactly as bntseq.c:144,173 does
 * (srand48(11), lrand48()&3) and recorded in .amb. */
int mc_index_build_fasta(const char *fasta_path, int32_t n_threads, mc_index **out);
/* Replaces bwa_idx_load() + RestoreReferenceInfo(). */
int mc_index_load(const char *prefix, mc_index **out);
/* Writes <prefix>.bwt/.sa/.pac/.ann/.amb byte-compatible with the reference's builder. */
int mc_index_save(const mc_index *idx, const char *prefix);
/* Wraps arrays the caller already holds (e.g. the reference's own bwaidx_t) without copying. */
int mc_index_wrap(const mc_index_view *view, mc_index **out);
int mc_index_get(const mc_index *idx, mc_index_view *view);
void mc_index_free(mc_index *idx);

/* ------------------------------------------------------------------------------------------------
 * Mapping context: one GPU, one full index replica in HBM, the device-resident pile-up profile and
 * the sequential state the reference keeps in globals (avgDist, totals; src/ReadMapping.cpp:20-21).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mc_ctx mc_ctx;

typedef struct {
    int32_t paired;            /* bPairEnd: reads arrive as adjacent mates (mate 1 first) */
    int32_t alg_ksw2;          /* !NW_ALG: 0 = nw_alignment, 1 = ksw2_alignment (-alg) */
    int32_t max_pos_diff;      /* MaxPosDiff (-indel), default 30 */
    int32_t max_clip;          /* MaxClipSize (-maxclip), default 5 */
    int32_t max_dup;           /* iMaxDuplicate (-dup), default 5, 1..15 */
    float   max_mismatch_rate; /* MaxMisMatchRate (-maxmm), default 0.05f */
    int32_t update_profile;    /* bVCFoutput: accumulate the pile-up profile */
    int32_t want_alignments;   /* bSAMoutput: return per-read candidates/fragments for SamReport */
    int32_t device;            /* CUDA device ordinal */
    int32_t shard_rank;        /* multi-GPU: this context's position in file order (0 when single) */
    int32_t shard_count;       /* multi-GPU: number of shards (1 when single) */
    int32_t reserved[5];       /* reserved[0] != 0: return the per-pair coordinates (mc_batch_out::pairs) even without want_alignments;
                                  reserved[1] != 0: keep the 128-row reference layout of the index in HBM (the path of texts >= 2^32 symbols);
                                  reserved[2] != 0: multi-GPU shards are independent libraries (no ordered exchange, see mc_comm_init) */
} mc_params;

void mc_params_default(mc_params *p);   /* defaults of reference src/main.cpp:159-191 */

int mc_ctx_create(const mc_index *idx, const mc_params *params, mc_ctx **out);
void mc_ctx_destroy(mc_ctx *ctx);

/* One batch = the NEXT n_reads reads of the library in file order (reference chunk protocol,
 * src/GetData.cpp:85-99): a multiple of MC_CHUNK_READS except for the last batch of a library. */
typedef struct {
    int64_t n_reads;
    const uint8_t *seq;       /* concatenated ASCII bases as read from FASTA/FASTQ (mate 2 NOT yet reversed) */
    const int64_t *seq_off;   /* n_reads + 1 offsets into seq */
} mc_batch_in;

/* AlnSummary_t (reference src/structure.h:135-140) + the slice of candidates of this read */
typedef struct { int32_t score, sub_score, best_idx, cand_begin, n_cand, rlen; } mc_read_out;
/* AlnCan_t (reference src/structure.h:125-133); frags only for score > 0 */
typedef struct { int32_t score, orientation, paired_idx, frag_begin, n_frag, pad; } mc_cand_out;
/* FragPair_t (reference src/structure.h:113-123); aln1 = aln[aln_off .. +aln_len), aln2 follows at aln_off+aln_cap */
typedef struct { int64_t gPos; int32_t rPos, rLen, gLen, bSimple, aln_off, aln_len, aln_cap, pad; } mc_frag_out;
/* CoordinatePair_t of GenCoordinatePair (reference src/ReadMapping.cpp:361-394), one per pair */
typedef struct { int64_t gPos1, gPos2, dist; } mc_pair_out;
/* what one 200-read chunk adds to the totals under OutputLock (reference src/ReadMapping.cpp:538) */
typedef struct { int32_t n_reads, mapped, paired, est_distance; int64_t dist_sum, len_sum; } mc_chunk_out;

typedef struct {
    int64_t n_reads, n_cands, n_frags, n_aln_bytes, n_pairs, n_chunks;
    const mc_read_out *reads;      /* n_reads (NULL unless want_alignments) */
    const mc_cand_out *cands;      /* arena; index through reads[].cand_begin */
    const mc_frag_out *frags;      /* arena; index through cands[].frag_begin */
    const uint8_t *aln;            /* arena of alignment strings */
    const mc_pair_out *pairs;      /* n_pairs (paired mode, with want_alignments or reserved[0]) */
    const mc_chunk_out *chunks;    /* n_chunks */
    int32_t replays;               /* chunks re-run because the avgDist speculation missed (diagnostic) */
} mc_batch_out;

/* Replaces the body of ReadMapping() (reference src/ReadMapping.cpp:416-646) for one batch:
 * seeding, clustering, pairing, rescue, gapped fills, scoring, pair statistics and the profile
 * update, bit-identical to a single reference thread processing the same reads in order.
 * Host batches of >= 400 k reads travel in four pieces and are seeded piece by piece while the rest is on the wire.
 * Every batch but the last of a library must be a whole number of 200-read chunks (MC_CHUNK_READS): the reference cuts the
 * library into chunks from its start, and a batch that ends inside one is taken as the end (a later batch is refused until
 * mc_reset). */
int mc_map_batch(mc_ctx *ctx, const mc_batch_in *in, mc_batch_out *out);

/* Page-locked host memory for batch inputs: reads placed here are DMA-ed to the GPU directly, pageable memory is
 * bounced through two pinned buffers (still correct, one extra host copy). */
int mc_host_alloc(size_t bytes, void **out);
void mc_host_free(void *p);

/* Sequential state (reference globals iTotalReadNum, iTotalMappingNum, iTotalPairedNum,
 * TotalPairedDistance, ReadLengthSum, avgDist; src/ReadMapping.cpp:20-21) */
typedef struct { int64_t total_reads, total_mapped, total_paired, total_distance, read_length_sum; uint32_t avg_dist; uint32_t pad; } mc_totals;
int mc_get_totals(const mc_ctx *ctx, mc_totals *out);
int mc_set_totals(mc_ctx *ctx, const mc_totals *in);
/* Back to the state of a fresh context (empty profile, avgDist = 1000, totals 0): the start of a new run. */
int mc_reset(mc_ctx *ctx);
/* The next batch starts a new library of the same run (the loop over -f files of Mapping(), reference
 * src/ReadMapping.cpp:705-748): the 200-read chunk grid restarts, while the profile, the totals and avgDist keep
 * accumulating.  Needed after a batch that was not a whole number of chunks (which ends a library). */
int mc_begin_library(mc_ctx *ctx);

/* ---- profile (reference MappingRecordArr / InsertSeqMap / DeleteSeqMap / BreakPointMap /
 *      InversionSiteVec / TranslocationSiteVec; src/structure.h:152-163,220-221) --------------- */

/* Packs the device counters of [beg, end) into MappingRecord_t records (16 bytes each: uint64
 * A:12,C:12,G:12,T:12,multi_hit:12,readCount:4 then uint16 F1,R2,F2,R1), saturating/wrapping as the
 * reference does. `out` must hold (end-beg)*16 bytes. */
int mc_profile_read(mc_ctx *ctx, int64_t beg, int64_t end, void *out);

/* CheckMappingCoverage / ReportDuplicationRate (reference src/ReadMapping.cpp:648-687) as device reductions over the
 * resident profile, without downloading it: columns with A+C+G+T > 0 (iAlignedBase) and their sum (iTotalCoverage);
 * columns with readCount > 0 and the sum of their readCount.  avgCov = (int)(coverage_sum / aligned_bases + .5),
 * duplication rate = 100 * (dup_reads - dup_sites) / dup_sites. */
typedef struct { int64_t aligned_bases, coverage_sum, dup_sites, dup_reads; } mc_profile_stats;
int mc_profile_summary(mc_ctx *ctx, mc_profile_stats *out);
/* Order-independent 128-bit fingerprint of the whole profile as mc_profile_read() would return it (every column's
 * MappingRecord_t mixed with its position, summed and xor-ed over the genome) - computed on the device; equal profiles
 * give equal fingerprints.  bench.py uses it to check that N GPUs on one library produce the single-GPU profile. */
int mc_profile_checksum(mc_ctx *ctx, uint64_t out[2]);

typedef struct { int64_t pos; int32_t kind /* 0 ins, 1 del */, len, count, seq_off; } mc_indel_rec;
/* Unique (pos, kind, sequence) triples with their uint16-wrapped counts, sorted by (kind,pos,seq). */
int mc_profile_indels(mc_ctx *ctx, const mc_indel_rec **recs, int64_t *n_recs, const uint8_t **seq_arena);
typedef struct { int64_t pos; int64_t count; } mc_breakpoint_rec;
int mc_profile_breakpoints(mc_ctx *ctx, const mc_breakpoint_rec **recs, int64_t *n_recs);
typedef struct { int64_t gPos, dist; } mc_site_rec;
/* kind 0 = InversionSiteVec, 1 = TranslocationSiteVec, sorted by gPos as the thread-end merge does */
int mc_profile_sites(mc_ctx *ctx, int32_t kind, const mc_site_rec **recs, int64_t *n_recs);

/* ---- SAM record fields of the batch just mapped (reference src/SamReport.cpp:7-316, 324-488): flag (SetPairedAlignmentFlag /
 *      SetSingledAlignmentFlag), MAPQ (EvaluateMAPQ), RNAME / POS (GetAlnCoordinate + DetermineCoordinate), CIGAR text
 *      (GenerateCIGARstring), mate position and TLEN, NM / AS / XS - computed on the device from the candidate, fragment and
 *      alignment-string arenas of the last mc_map_batch / mc_map_staged call, one record per read (the reference's default
 *      bUnique = true).  The caller prints the line; it owns QNAME, SEQ and QUAL:
 *        QNAME flag RNAME pos mapq CIGAR (has_mate ? "=" mate_pos tlen : "*" 0 0) SEQ QUAL [NM:i:nm] AS:i:as XS:i:xs
 *      with SEQ = the read as it stands in the FASTQ file (both mates), reverse-complemented and QUAL reversed when
 *      `reverse` is set.  Unmapped read: chrom = -1 ("*", pos 0, mapq 0,
 *      CIGAR "*"), nm = -1 (no NM tag).  flag = -1: the reference prints no line for this read. */
typedef struct {
	int64_t pos, mate_pos;
	int32_t flag, chrom, mapq, tlen, nm, as, xs, cigar_off, cigar_len, reverse, has_mate, pad;
} mc_sam_rec;
/* `recs` (one per read of the batch) and `cigar_arena` stay valid until the next call on ctx. */
int mc_sam_records(mc_ctx *ctx, const mc_sam_rec **recs, int64_t *n_recs, const uint8_t **cigar_arena);

/* The SAM TEXT of the batch just mapped, assembled on the device: the lines GeneratePairedSamStream / GenerateSingleSamStream
 * print (reference src/SamReport.cpp:324-488), in read order, each terminated by '\n' (no line for a read the reference
 * prints none for).  The batch must have been staged by mc_ingest_fastq(slot) and mapped by mc_map_staged(slot): its FASTQ
 * text is still in the slot, and QNAME (IdentifyHeaderBegPos / IdentifyHeaderEndPos, src/GetData.cpp:3-21), SEQ and QUAL
 * are taken from there.  all_best != 0 is the reference's -m (bUnique = false): one line per candidate that reaches the
 * read's best score instead of one line per read.  `text` is page-locked library memory, valid until the next call. */
int mc_sam_text(mc_ctx *ctx, int32_t slot, int32_t all_best, const uint8_t **text, int64_t *n_bytes);

/* ---- variant-calling scan (reference src/VariantCalling.cpp:106-120 CalBlockReadDepth, :550-680 IdentifyVariants,
 *      :60-98 GetAreaIndFrequency, :523-548 DetermineGenotype) over the device-resident profile, without downloading
 *      the 16-byte-per-column MappingRecordArr.  The columns are scanned on the GPU in blocks of 100 (the reference's
 *      BlockSize): block depths, gap / duplication runs and gVCF segments are carried across blocks with device scans.
 *      The +-5-column indel-window join runs over the aggregated indel records of mc_profile_indels() (sparse, host).
 *      Output = the reference's VariantVec after IdentifyVariants (+ RemoveConsecutiveGenomicVariant under gvcf), sorted
 *      by (gPos, VarType) like CompByVarPos; structural variants (IdentifyInversions / IdentifyTranslocations) and the
 *      VCF text stay with the caller, who has everything they need in `record` and `block_depth`.
 *      Fields the reference leaves stale for a type (e.g. AD_alt of a gap record) are 0 here.  DP of gap / dup records is
 *      the run length truncated to 16 bits exactly as Variant_t::DP does. */
typedef struct {
	int32_t min_allele_depth;   /* MinAlleleDepth (-ad), >= 1 */
	float frequency_thr;        /* FrequencyThr (-maxmm is unrelated; src/main.cpp:183) */
	int32_t somatic, gvcf, monomorphic, ploidy, min_cnv_size, min_unmapped_size;
} mc_vc_params;
void mc_vc_params_default(mc_vc_params *p);   /* src/main.cpp:159-185 */
enum { MC_VAR_SUB = 0, MC_VAR_INS = 1, MC_VAR_DEL = 2, MC_VAR_CNV = 5, MC_VAR_UMR = 6, MC_VAR_NOR = 10, MC_VAR_MON = 11 };
typedef struct {
	int64_t gPos;
	uint64_t record[2];         /* MappingRecord_t of gPos (as mc_profile_read packs it); 0 for CNV / UMR */
	int32_t alt_off, alt_len;   /* INS / DEL: ALTstr in the sequence arena of mc_profile_indels(); otherwise 0, 0 */
	uint16_t DP, AD_ref, AD_alt;
	uint8_t GenoType, qscore, VarType;
	char alt[3];                /* SUB: "X" or "X,Y" (unused bytes 0) */
} mc_variant_rec;
/* `recs`, `alt_arena` and `block_depth` stay valid until the next mc_variant_scan / mc_profile_indels / mc_reset on ctx. */
int mc_variant_scan(mc_ctx *ctx, const mc_vc_params *vp, const mc_variant_rec **recs, int64_t *n_recs,
                    const uint8_t **alt_arena, const int32_t **block_depth, int64_t *n_blocks);

/* Multi-GPU (one context per GPU, one process per GPU): reads shard across the ranks, every rank holds a full index
 * replica.  mc_comm_unique_id() on rank 0 -> ship the 128 bytes to the other ranks -> mc_comm_init() on every rank.
 *
 * From then on the ranks work on ONE library and mc_map_batch() is a collective: within a call, the reads of rank r follow
 * those of rank r-1 in file order, and the next call continues after the last rank.  The three things the reference does
 * sequentially in file order - the avgDist feedback (src/ReadMapping.cpp:538-539), the PCR-duplicate gate
 * (src/AlignmentProfile.cpp:76-77) and the thread-local discordant-pair state (src/ReadMapping.cpp:486-522) - are exchanged
 * over NCCL inside the call, so every record a rank returns, and the totals (the same on every rank), equal what a single
 * reference thread produces for the whole library.  Every rank must pass at least one chunk per call.
 * mc_params.reserved[2] = 1 switches the exchange off: every shard is then mapped as a library of its own.
 *
 * mc_profile_allreduce() sums the device counters of all ranks (ncclAllReduce over NVLink) and gathers the indel /
 * break-point / SV-site records device to device, so that every rank holds the whole-library profile; it ends a library
 * run (mapping more reads afterwards would count the other ranks' share again).  `nccl_comm` may be an existing
 * ncclComm_t of the caller (independent shards only), or NULL to use the communicator of mc_comm_init(). */
int mc_comm_unique_id(uint8_t *out128);
int mc_comm_init(mc_ctx *ctx, const uint8_t *id128, int32_t rank, int32_t n_ranks);
int mc_profile_allreduce(mc_ctx *ctx, void *nccl_comm);
/* The scalable variant for genomes of gigabases ("ncclReduceScatter ... so each GPU finalises 1/N of the
 * genome"): every rank ends up with the library's counters of ONE tile of the genome (equal tiles of whole 25600-column
 * units, rank r owns tile r; ncclReduceScatter in place - half the NVLink traffic of the all-reduce and no 33 B x G image
 * per rank to finish) and the read-out calls that follow become collectives over the tiles, to be made by every rank:
 *   mc_variant_scan       scans the own tile (run carriers of the earlier tiles come in through one small all-gather) and
 *                         gathers the records and block depths of all ranks: every rank returns the whole genome's result
 *                         (a gVCF scan needs mc_profile_allreduce: its runs look ahead across tiles)
 *   mc_profile_summary, mc_profile_checksum   partial results of the tiles, combined
 *   mc_profile_read       serves columns of the own tile only (mc_profile_owned tells which)
 * Indel, break-point and SV-site records are gathered to every rank as by mc_profile_allreduce.  Ends the library run. */
int mc_profile_reduce_scatter(mc_ctx *ctx);
int mc_profile_owned(const mc_ctx *ctx, int64_t *beg, int64_t *end);   /* [0, genome) unless reduce-scattered */

/* ---- read ingest (first "next" row) -----------------------------------------------------------
 * GetNextEntry / GetNextChunk, FASTQ branch (reference src/GetData.cpp:32-99), on the device: the caller hands over raw
 * blocks of uncompressed FASTQ text (one per mate file, or one with the mates as adjacent records) and gets a batch staged
 * in device slot `slot`, ready for mc_map_staged().  A record is four lines; the read is the second and its length is the
 * length of that line without the newline.  Only whole records are consumed; unless `final_block` is set the batch is cut
 * to a multiple of the reference's 200-read chunk, so that batch boundaries do not move its chunk grid.  `consumed1/2`
 * tell the caller where the next block has to start.  Read names and qualities stay with the caller (the mapping path
 * does not use them; a SAM writer needs them). */
typedef struct {
    const uint8_t *text1; int64_t len1;   /* block of the first (or only) FASTQ file */
    const uint8_t *text2; int64_t len2;   /* block of the second mate file, or NULL */
    int64_t max_reads;                    /* 0 = as many as the blocks hold */
    int32_t final_block;                  /* the blocks reach the end of the file(s) */
    int32_t format;                       /* 0 = FASTQ (four lines per record); 1 = FASTA reads with the sequence on ONE line (two
                                             lines per record: '>' header, bases - the FASTA branch of GetNextEntry, src/GetData.cpp:54-76,
                                             for records that are not wrapped; a wrapped record makes the call fail with MC_ERR_ARG) */
} mc_fastq_in;
typedef struct { int64_t n_reads, consumed1, consumed2, n_bases;
                 int64_t records1, records2;   /* whole records the blocks held (before the cut to n_reads) */ } mc_fastq_out;
int mc_ingest_fastq(mc_ctx *ctx, const mc_fastq_in *in, int32_t slot, mc_fastq_out *out);
/* Optional: queues the host -> device copy of the blocks a later mc_ingest_fastq(ctx, in, slot, ..) is going to parse (same
 * pointers and lengths) and returns at once; the copy is plain DMA on a stream of its own, so PCIe stays busy while an
 * earlier block is parsed and a still earlier one is mapped.  The blocks must lie in page-locked memory (mc_host_alloc)
 * and stay untouched until that mc_ingest_fastq has returned. */
int mc_ingest_prefetch(mc_ctx *ctx, const mc_fastq_in *in, int32_t slot);
/* mc_ingest_fastq() works on the context's copy stream with scratch of its own: a second host thread may ingest the next
 * block into slot s' while mc_map_staged() maps slot s != s' (the one exception to "calls on one ctx are serialised";
 * the text must then lie in page-locked memory, mc_host_alloc). */

/* ---- operator-level entry points (per-kernel parity tests and micro-benchmarks) ---------------- */

/* BWT_Search (reference src/bwt_search.cpp:121) for n independent (read, start) queries.
 * codes: concatenated 0..4 codes, off[n+1]; start[n].  out_len/out_freq[n]; out_loc[n*MC_MAX_OCC]. */
int mc_bwt_search_batch(mc_ctx *ctx, int64_t n, const uint8_t *codes, const int64_t *off, const int32_t *start,
                        int32_t *out_len, int32_t *out_freq, uint64_t *out_loc);
/* ProduceReadAlignment in its setting (reference src/structure.h:242, src/ReadAlignment.cpp:306-430): SimplePairClustering ->
 * RemoveRedundantAlnCan -> ProduceReadAlignment for n independent reads, the single-end branch of ReadMapping()
 * (src/ReadMapping.cpp:575-585).  `out` as mc_map_batch with want_alignments (candidates, fragment lists, alignment strings,
 * scores, AlnSummary); the context's totals, avgDist, chunk grid and profile are untouched.  Mates are not reversed. */
int mc_read_alignment_batch(mc_ctx *ctx, const mc_batch_in *in, mc_batch_out *out);
/* AlignmentRescue in its setting (reference src/structure.h:245, src/AlignmentRescue.cpp:28-111): the pairs of the batch through
 * seeding, clustering, pairing / masking, AlignmentRescue(EstiDistance, read1, read2) and ProduceReadAlignment with
 * EstiDistance = (int)(avg_dist * 1.5) for every pair (src/ReadMapping.cpp:462): no avgDist feedback between chunks, the
 * context's sequential state is untouched, so a pair's records depend on the pair and avg_dist alone. */
int mc_rescue_batch(mc_ctx *ctx, const mc_batch_in *in, uint32_t avg_dist, mc_batch_out *out);
/* UpdateProfile / UpdateMultiHitCount as a step of its own (reference src/structure.h:248-249, src/AlignmentProfile.cpp:41-271,
 * call site src/ReadMapping.cpp:559-568).  After mc_defer_profile(ctx, 1) a mapped batch leaves the profile untouched and its
 * candidates, fragment lists and alignment strings in the device arenas; mc_update_profile_last(ctx) applies the update to
 * exactly those reads (dedup gate, strand / base / multi-hit counters, indel and break-point records), once, before the
 * next batch is mapped.  Not with the ordered multi-GPU exchange. */
int mc_defer_profile(mc_ctx *ctx, int32_t on);
int mc_update_profile_last(mc_ctx *ctx);
/* IdentifySimplePairs + SimplePairClustering (reference src/ReadMapping.cpp:125-226) for n independent reads, taken as they
 * are (no mate reversal): per read its simple pairs sorted by (PosDiff, rPos) - the sentinel left out - and its candidate
 * clusters as [pair_begin, pair_end) slices of that list (for a tandem-repeat cluster: the best equal-PosDiff run).
 * The arrays are the library's, valid until the next call on the context. */
typedef struct { int64_t gPos; int32_t rPos, len; } mc_simple_pair;
typedef struct { int32_t score, pair_begin, pair_end; } mc_cluster;
typedef struct {
    int64_t n_reads;
    const int64_t *pair_off;        /* n_reads + 1: read r owns pairs [pair_off[r], pair_off[r] + n_pairs[r]) */
    const int32_t *n_pairs;         /* n_reads */
    const mc_simple_pair *pairs;
    const int64_t *cluster_off;     /* n_reads + 1 (same spacing as pair_off) */
    const int32_t *n_clusters;      /* n_reads */
    const mc_cluster *clusters;     /* pair_begin / pair_end index `pairs` directly */
} mc_seed_cluster_out;
int mc_seed_cluster_batch(mc_ctx *ctx, const mc_batch_in *in, mc_seed_cluster_out *out);
/* nw_alignment / ksw2_alignment (reference src/nw_alignment.cpp:18, src/ksw2_alignment.cpp:250) for n
 * independent problems.  s1/s2 concatenated ASCII with offsets; out1/out2 receive the gapped strings at
 * out_off[i] (caller provides out_off[i+1]-out_off[i] >= len1+len2); out_len[i] = aligned length. */
int mc_align_batch(mc_ctx *ctx, int32_t use_ksw2, int64_t n, const uint8_t *s1, const int64_t *off1, const uint8_t *s2,
                   const int64_t *off2, const int64_t *out_off, uint8_t *out1, uint8_t *out2, int32_t *out_len);

/* ---- measurement hooks ---------------------------------------------------------------------- */
typedef struct {
    double ms_seed, ms_locate, ms_cluster, ms_pair, ms_align, ms_profile, ms_h2d, ms_d2h, ms_total;
    int64_t seed_blocks;      /* 64-byte occ blocks the reference algorithm touches while extending */
    int64_t locate_blocks;    /* 64-byte blocks touched by the LF walks of bwt_sa */
    int64_t sa_reads;         /* 8-byte sampled-SA reads */
    int64_t dp_cells;         /* sum of m*n over the gapped fills */
    int64_t dp_tasks;
    int64_t profile_columns;  /* MappingRecord_t columns the reference's pile-up update touches */
    int64_t profile_atomics;  /* atomic updates the device profile actually needed for them */
    int64_t kernel_launches;
    double ms_reduce;         /* device time of mc_profile_allreduce (multi-GPU), CUDA events on the context's stream */
    int64_t h2d_bytes;        /* bytes this process copied host -> device / device -> host through the library since the last */
    int64_t d2h_bytes;        /* mc_reset_stats (counted where the copies are issued) */
    double ms_reset;          /* device time of mc_reset (zeroing the profile); also part of ms_total */
} mc_stats;
int mc_get_stats(const mc_ctx *ctx, mc_stats *out);   /* accumulated since create / last reset */
int mc_reset_stats(mc_ctx *ctx);
/* Double-buffered feed: a batch is copied into one of eight device slots and mapped from there, any number of times.
 * mc_stage_batch_async() only queues the copy (and the reversal of mate 2) on the context's copy stream and returns; the
 * caller's buffers must stay untouched until the next mc_map_staged() of that slot has returned (page-locked buffers from
 * mc_host_alloc are copied by DMA, pageable ones are bounced synchronously).  A host that maps batch i from slot i & 1
 * after staging batch i + 1 into the other slot hides the PCIe transfer behind the mapping of the previous batch; this is
 * the loop bench.py times as `e2e`.  mc_stage_batch() is the same followed by a wait.  Do not stage into a slot while it
 * is being mapped. */
int mc_stage_batch_async(mc_ctx *ctx, const mc_batch_in *in, int32_t slot);
int mc_stage_batch(mc_ctx *ctx, const mc_batch_in *in, int32_t slot);
int mc_map_staged(mc_ctx *ctx, int32_t slot, mc_batch_out *out);

#ifdef __cplusplus
}
#endif
#endif /* MAPCALLER_B200_H */
