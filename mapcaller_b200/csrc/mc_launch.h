// Kernel wrappers: every stage body of mc_stages.h becomes one __global__ kernel (one thread per work
// item, 256-thread blocks) plus a launch_<stage>() function.  Under MC_HOSTEMU (developer harness, see
// mc_device.h) the launchers are plain loops.
#ifndef MC_LAUNCH_H
#define MC_LAUNCH_H

#include "mc_stages.h"

#ifdef MC_HOSTEMU
#include <algorithm>
typedef int mc_stream_t;
#define MC_LAUNCH1(name) \
	static void launch_##name(const PipeArgs& a, int64_t n, mc_stream_t) { for (int64_t i = 0; i < n; i++) name##_body(i, a); }
#define MC_LAUNCH2(name) \
	static void launch_##name(const PipeArgs& a, const ProfArgs& q, int64_t n, mc_stream_t) { for (int64_t i = 0; i < n; i++) name##_body(i, a, q); }
// bodies that take part in block-wide cursor bumps get a `live` flag (false for the padding threads of the last block)
#define MC_LAUNCH1B(name) \
	static void launch_##name(const PipeArgs& a, int64_t n, mc_stream_t) { for (int64_t i = 0; i < n; i++) name##_body(i, true, a); }
#define MC_LAUNCH2B(name) \
	static void launch_##name(const PipeArgs& a, const ProfArgs& q, int64_t n, mc_stream_t) { for (int64_t i = 0; i < n; i++) name##_body(i, true, a, q); }
static void launch_rescue(const PipeArgs& a, int64_t n, mc_stream_t)
{
	for (int64_t i = 0; i < n; i++) rwenum_body(i, a);
	if (a.st->overflow & 0xFF) return;
	const int64_t nw = (int64_t)*a.rwin_bump - a.rwin_begin;
	for (int64_t i = 0; i < nw; i++) rwin_body(i, 0, 1, a, 0, 0);
	for (int64_t i = 0; i < n; i++) rcommit_body(i, a);
}
static void launch_locate(const PipeArgs& a, int64_t n, mc_stream_t) { if (n > 0) locate_body(0, 1, a); }
static void launch_prep(const PipeArgs& a, int64_t first, int64_t n, mc_stream_t) { for (int64_t r = first | 1; r < n; r += 2) prep_body(r, 0, 1, a); }
static void launch_seed(const PipeArgs& a, int64_t first, int64_t n, mc_stream_t) { seed_body(0, 1, first, n, a, nullptr); }
static void launch_piece(const PipeArgs& a, int64_t, mc_stream_t) { const int64_t n = (int64_t)*a.ptask_bump - a.ptask_begin; for (int64_t i = 0; i < n; i++) piece_body(i, 0, 1, a); }
static void launch_chunkstat(const PipeArgs& a, int64_t n, mc_stream_t) { for (int64_t i = 0; i < n; i++) chunkstat_body(i, 0, 1, a); }
static void launch_profpiece(const PipeArgs& a, const ProfArgs& q, int64_t, mc_stream_t) { const int64_t n = (int64_t)*a.ptask_bump; for (int64_t i = 0; i < n; i++) a.st->profile_atomics += profpiece_body(i, 0, 1, a, q); }
static void launch_disclist(const PipeArgs& a, int64_t n, DiscRec* out, mc_u64* bump, int64_t cap, mc_stream_t) { for (int64_t i = 0; i < n; i++) disclist_body(i, a, out, bump, cap); }
static void launch_profsum(const DevProfile& p, int64_t G, int64_t nb, int64_t b0, int64_t b1, int64_t* sums, mc_stream_t) { for (int64_t b = b0; b < b1; b++) profsum_body(b, 0, 1, p, G, nb, sums); }
static void launch_profpack(const DevIndex& ix, const DevProfile& p, int64_t nb, const int64_t* pre, int64_t b0, int64_t b1, int64_t beg, int64_t end, uint64_t* out, mc_stream_t)
{ for (int64_t b = b0; b < b1; b++) profpack_body(b, 0, 1, ix, p, nb, pre, beg, end, out); }
static void launch_cbwt_build(int64_t n, const uint32_t* src, uint32_t* dst, mc_stream_t) { for (int64_t b = 0; b < n; b++) mc_cbwt_build_body(b, src, dst); }
static void launch_sa_dense(const DevIndex& ix, int64_t n, int shift, uint32_t* o32, uint64_t* o64, mc_stream_t) { for (int64_t j = 0; j < n; j++) mc_sa_dense_body(j, ix, shift, o32, o64); }
static void launch_ktab_build(const DevIndex& ix, int k, uint32_t* o32, uint64_t* o64, mc_stream_t) { for (int64_t m = 0; m < (1ll << (2 * k)); m++) mc_ktab_build_body(m, ix, k, o32, o64); }
static void launch_profstat(int64_t n, const uint64_t* recs, mc_u64* acc, mc_stream_t) { for (int64_t i = 0; i < n; i++) profstat_body(i, recs, acc); }
static void launch_indkey(int64_t n, const mc_indel_rec* recs, const uint8_t* seq, uint64_t* keys, uint32_t* idx, mc_stream_t) { for (int64_t i = 0; i < n; i++) indkey_body(i, recs, seq, keys, idx); }
static void launch_indlen(int64_t n, const mc_indel_rec* recs, const uint32_t* idx, uint32_t* len, mc_stream_t) { for (int64_t j = 0; j < n; j++) indlen_body(j, recs, idx, len); }
static void launch_indgather(int64_t n, const mc_indel_rec* recs, const uint8_t* seq, const uint32_t* idx, const int64_t* off, mc_indel_rec* out, uint8_t* out_seq, mc_stream_t) { for (int64_t j = 0; j < n; j++) indgather_body(j, recs, seq, idx, off, out, out_seq); }
static void launch_profhash(int64_t g0, int64_t n, const uint64_t* recs, mc_u64* acc, mc_stream_t) { for (int64_t i = 0; i < n; i++) { const uint64_t h = profhash_of(g0 + i, recs, i); acc[0] += h; acc[1] ^= h; } }
static void launch_gatecnt(const PipeArgs& a, const ProfArgs& q, int64_t n, uint64_t* list, mc_u64* bump, mc_stream_t) { for (int64_t i = 0; i < n; i++) gatecnt_body(i, a, q, list, bump); }
static void launch_gateadd(const PipeArgs& a, int64_t n, const uint64_t* list, mc_stream_t) { for (int64_t i = 0; i < n; i++) gateadd_body(i, a, list); }
static void launch_gatedense_fill(const PipeArgs& a, const ProfArgs& q, int64_t n, uint8_t* dense, mc_stream_t) { for (int64_t i = 0; i < n; i++) gatedense_fill_body(i, a, q, dense); }
static void launch_gatedense_apply(const PipeArgs& a, int64_t G, const uint8_t* all, size_t pitch, int r0, int r1, mc_stream_t) { for (int64_t g = 0; g < G; g++) gatedense_apply_body(g, a, all, pitch, r0, r1); }
static void launch_fqcount(const FastqArgs& q, int f, int64_t n, mc_stream_t) { for (int64_t t = 0; t < n; t++) fqcount_body(t, f, q); }
static void launch_fqlines(const FastqArgs& q, int f, int64_t n, mc_stream_t) { for (int64_t t = 0; t < n; t++) fqlines_body(t, f, q); }
static void launch_fqread(const FastqArgs& q, int64_t n, mc_stream_t) { for (int64_t r = 0; r < n; r++) fqread_body(r, q); }
static void launch_fqcopy(const FastqArgs& q, int64_t n, mc_stream_t) { for (int64_t r = 0; r < n; r++) fqcopy_body(r, 0, 1, q); }
static void launch_bwtsearch(const SearchArgs& a, int64_t n, mc_stream_t) { for (int64_t q = 0; q < n; q++) bwtsearch_body(q, a); }
static void launch_vcdepth(const VcArgs& a, int64_t b0, int64_t b1, mc_stream_t) { for (int64_t b = b0; b < b1; b++) vcdepth_body(b, a); }
static void launch_vcscan(const VcArgs& a, int64_t b0, int64_t b1, bool emit, mc_stream_t) { for (int64_t b = b0; b < b1; b++) vcscan_body(b, a, emit); }
static void launch_vckey(int64_t n, const mc_variant_rec* recs, uint64_t* keys, uint32_t* idx, mc_u64* n_valid, mc_stream_t) { for (int64_t i = 0; i < n; i++) vckey_body(i, recs, keys, idx, n_valid); }
static void launch_vcgather(int64_t n, const mc_variant_rec* in, const uint32_t* idx, mc_variant_rec* out, mc_stream_t) { for (int64_t j = 0; j < n; j++) vcgather_body(j, in, idx, out); }
static void launch_samrec(const SamArgs& a, int64_t n, bool emit, mc_stream_t) { for (int64_t r = 0; r < n; r++) samrec_body(r, a, emit); }
static void launch_samtext(const SamTextArgs& t, int64_t n, bool emit, mc_stream_t) { for (int64_t r = 0; r < n; r++) samtext_body(r, t, emit); }
static void device_incmax_i64(int64_t* a, int64_t n, void*, mc_stream_t) { for (int64_t i = 1; i < n; i++) if (a[i] < a[i - 1]) a[i] = a[i - 1]; }
static void device_exscan_i64(int64_t* a, int64_t n, int64_t* total, void*, mc_stream_t) { int64_t s = 0; for (int64_t i = 0; i < n; i++) { int64_t v = a[i]; a[i] = s; s += v; } *total = s; }
#include <vector>
static int64_t g_launches = 0;
#define MC_SLOTS 8
#else
#include <atomic>
typedef cudaStream_t mc_stream_t;
static std::atomic<int64_t> g_launches(0);   // a second host thread may be staging the next batch (mc_ingest_fastq)
#define MC_SLOTS 8   // device slots for staged batches
#define MC_BLOCK 256
#define MC_LAUNCH1(name) \
	__global__ void __launch_bounds__(MC_BLOCK) mc_##name##_kernel(const PipeArgs a, int64_t n) \
	{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) name##_body(i, a); } \
	static void launch_##name(const PipeArgs& a, int64_t n, mc_stream_t s) \
	{ if (n > 0) { mc_##name##_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, n); g_launches++; } }
// bodies that take part in block-wide cursor bumps: every thread of the block runs the body, `live` = inside the batch
#define MC_LAUNCH1B(name) \
	__global__ void __launch_bounds__(MC_BLOCK) mc_##name##_kernel(const PipeArgs a, int64_t n) \
	{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; name##_body(i, i < n, a); } \
	static void launch_##name(const PipeArgs& a, int64_t n, mc_stream_t s) \
	{ if (n > 0) { mc_##name##_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, n); g_launches++; } }
#define MC_LAUNCH2B(name) \
	__global__ void __launch_bounds__(MC_BLOCK) mc_##name##_kernel(const PipeArgs a, const ProfArgs q, int64_t n) \
	{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; name##_body(i, i < n, a, q); } \
	static void launch_##name(const PipeArgs& a, const ProfArgs& q, int64_t n, mc_stream_t s) \
	{ if (n > 0) { mc_##name##_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, q, n); g_launches++; } }
#define MC_LAUNCH2(name) \
	__global__ void __launch_bounds__(MC_BLOCK) mc_##name##_kernel(const PipeArgs a, const ProfArgs q, int64_t n) \
	{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) name##_body(i, a, q); } \
	static void launch_##name(const PipeArgs& a, const ProfArgs& q, int64_t n, mc_stream_t s) \
	{ if (n > 0) { mc_##name##_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, q, n); g_launches++; } }
// persistent lanes: a fixed grid, every thread takes locations tid, tid + nthreads, ...
__global__ void __launch_bounds__(MC_BLOCK) mc_locate_kernel(const PipeArgs a)
{ locate_body(blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, a); }
static void launch_locate(const PipeArgs& a, int64_t n, mc_stream_t s)
{
	if (n <= 0) return;
	int64_t blocks = (n + MC_BLOCK - 1) / MC_BLOCK; if (blocks > 148 * 8) blocks = 148 * 8;   // 8 resident 256-thread blocks per SM
	mc_locate_kernel<<<(unsigned)blocks, MC_BLOCK, 0, s>>>(a); g_launches++;
}
__global__ void __launch_bounds__(MC_BLOCK) mc_disclist_kernel(const PipeArgs a, int64_t n, DiscRec* out, mc_u64* bump, int64_t cap)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) disclist_body(i, a, out, bump, cap); }
static void launch_disclist(const PipeArgs& a, int64_t n, DiscRec* out, mc_u64* bump, int64_t cap, mc_stream_t s)
{ if (n > 0) { mc_disclist_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, n, out, bump, cap); g_launches++; } }
// one warp per 200-read chunk
__global__ void __launch_bounds__(MC_BLOCK) mc_chunkstat_kernel(const PipeArgs a, int64_t n)
{
	const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
	if (w < n) chunkstat_body(w, threadIdx.x & 31, 32, a);
}
static void launch_chunkstat(const PipeArgs& a, int64_t n, mc_stream_t s)
{ if (n > 0) { mc_chunkstat_kernel<<<(unsigned)((n * 32 + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, n); g_launches++; } }
// normal pieces: persistent 8-lane tiles over the current attempt's piece list (four independent load chains per warp)
__global__ void __launch_bounds__(MC_BLOCK) mc_piece_kernel(const PipeArgs a)
{
	const int64_t tile = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3, n_tiles = ((int64_t)gridDim.x * blockDim.x) >> 3;
	const int64_t n = (int64_t)*a.ptask_bump - a.ptask_begin;
	for (int64_t t = tile; t < n; t += n_tiles) piece_body(t, threadIdx.x & 7, 8, a);
}
static void launch_piece(const PipeArgs& a, int64_t max_tasks, mc_stream_t s)
{
	if (max_tasks <= 0) return;
	int64_t blocks = (max_tasks * 8 + MC_BLOCK - 1) / MC_BLOCK; if (blocks > 148 * 8) blocks = 148 * 8;
	mc_piece_kernel<<<(unsigned)blocks, MC_BLOCK, 0, s>>>(a); g_launches++;
}
// mate reversal and seeding take a read range [first, n): a batch that arrives from the host in pieces is seeded piece by
// piece while the next piece is still on the wire (mc_map_batch)
__global__ void __launch_bounds__(MC_BLOCK) mc_prep_kernel(const PipeArgs a, int64_t first_pair, int64_t n_pairs)
{
	const int64_t w = first_pair + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
	if (w < n_pairs) prep_body(2 * w + 1, threadIdx.x & 31, 32, a);
}
static void launch_prep(const PipeArgs& a, int64_t first, int64_t n, mc_stream_t s)
{
	const int64_t p0 = first / 2, p1 = n / 2;
	if (!a.pr.paired || p1 <= p0) return;
	mc_prep_kernel<<<(unsigned)(((p1 - p0) * 32 + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, p0, p1); g_launches++;
}
// persistent lanes (seed_walk): a fixed grid of MINB resident blocks per SM
template <int MINB> __global__ void __launch_bounds__(MC_BLOCK, MINB) mc_seed_kernel(const PipeArgs a, int64_t first, int64_t n)
{
	__shared__ uint64_t rows[MC_BLOCK * MC_SEED_ROW_WORDS];          // a row per lane for the read it is working on (mc_stages.h)
	seed_body(blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, first, n, a, rows + threadIdx.x * MC_SEED_ROW_WORDS);
}
static void launch_seed(const PipeArgs& a, int64_t first, int64_t n, mc_stream_t s)
{
	if (n <= first) return;
	static const int forced = getenv("MC_SEED_MINB") ? atoi(getenv("MC_SEED_MINB")) : 0;
	const int variant = forced ? forced : 4;
	int64_t g = (n - first + MC_BLOCK - 1) / MC_BLOCK; if (g > 148 * variant) g = 148 * variant;
	if (variant == 6) mc_seed_kernel<6><<<(unsigned)g, MC_BLOCK, 0, s>>>(a, first, n);
	else if (variant == 3) mc_seed_kernel<3><<<(unsigned)g, MC_BLOCK, 0, s>>>(a, first, n);
	else if (variant == 5) mc_seed_kernel<5><<<(unsigned)g, MC_BLOCK, 0, s>>>(a, first, n);
	else mc_seed_kernel<4><<<(unsigned)g, MC_BLOCK, 0, s>>>(a, first, n);
	g_launches++;
}
// rescue (mc_stages_pair.h): enumerate the windows of the attempt's rescue pairs, search them with persistent thread blocks
// (block b takes windows b, b + n_blocks, ...), commit per pair.  A window search is a chain of short loops over small
// tables (word list and its hash chains, diagonal histogram, staged window): the tables live in shared memory (32 KB per block - enough for the 1500-base windows of the warm-up
// chunks at 150-base reads -, six blocks per SM) and the 128 threads of the block share every loop - windows in repeats cost 100x the typical one and
// their latency is what a replay attempt waits for.
#define MC_RESCUE_THREADS 128
#define MC_RESCUE_SMEM (32 * 1024)
__global__ void __launch_bounds__(MC_BLOCK) mc_rwenum_kernel(const PipeArgs a)
{
	const int64_t n = (int64_t)*a.rtask_bump - a.rtask_begin;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) rwenum_body(t, a);
}
__global__ void __launch_bounds__(MC_RESCUE_THREADS, 6) mc_rescue_kernel(const PipeArgs a)
{
	extern __shared__ __align__(16) uint8_t rescue_smem[];
	if (a.st->overflow & 0xFF) return;          // the window list is incomplete: the attempt is going to be repeated with larger arenas
	const int64_t n = (int64_t)*a.rwin_bump - a.rwin_begin;
	for (int64_t t = blockIdx.x; t < n; t += gridDim.x) { rwin_body(t, threadIdx.x, MC_RESCUE_THREADS, a, rescue_smem, MC_RESCUE_SMEM); __syncthreads(); }
}
__global__ void __launch_bounds__(MC_BLOCK) mc_rcommit_kernel(const PipeArgs a)
{
	if (a.st->overflow & 0xFF) return;
	const int64_t n = (int64_t)*a.rtask_bump - a.rtask_begin;
	for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) rcommit_body(t, a);
}
static void launch_rescue(const PipeArgs& a, int64_t max_tasks, mc_stream_t s)
{
	if (max_tasks <= 0) return;
	static bool configured[64];          // the attribute is per device: a process may hold contexts on several GPUs
	int dev = 0; cudaGetDevice(&dev); dev &= 63;
	if (!configured[dev]) { cudaFuncSetAttribute(mc_rescue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MC_RESCUE_SMEM); configured[dev] = true; }
	int64_t tb = (max_tasks + MC_BLOCK - 1) / MC_BLOCK; if (tb > 148 * 2) tb = 148 * 2;
	mc_rwenum_kernel<<<(unsigned)tb, MC_BLOCK, 0, s>>>(a); g_launches++;
	mc_rescue_kernel<<<148 * 6, MC_RESCUE_THREADS, MC_RESCUE_SMEM, s>>>(a); g_launches++;
	mc_rcommit_kernel<<<(unsigned)tb, MC_BLOCK, 0, s>>>(a); g_launches++;
}
__global__ void __launch_bounds__(MC_BLOCK) mc_profstat_kernel(int64_t n, const uint64_t* recs, mc_u64* acc)
{ for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ((n + 31) & ~31ll); i += (int64_t)gridDim.x * blockDim.x) if (i < n) profstat_body(i, recs, acc); }
static void launch_profstat(int64_t n, const uint64_t* recs, mc_u64* acc, mc_stream_t s)
{ if (n > 0) { int64_t b = (n + MC_BLOCK - 1) / MC_BLOCK; if (b > 148 * 8) b = 148 * 8; mc_profstat_kernel<<<(unsigned)b, MC_BLOCK, 0, s>>>(n, recs, acc); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_indkey_kernel(int64_t n, const mc_indel_rec* recs, const uint8_t* seq, uint64_t* keys, uint32_t* idx)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) indkey_body(i, recs, seq, keys, idx); }
__global__ void __launch_bounds__(MC_BLOCK) mc_indlen_kernel(int64_t n, const mc_indel_rec* recs, const uint32_t* idx, uint32_t* len)
{ int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (j < n) indlen_body(j, recs, idx, len); }
__global__ void __launch_bounds__(MC_BLOCK) mc_indgather_kernel(int64_t n, const mc_indel_rec* recs, const uint8_t* seq, const uint32_t* idx, const int64_t* off, mc_indel_rec* out, uint8_t* out_seq)
{ int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (j < n) indgather_body(j, recs, seq, idx, off, out, out_seq); }
static void launch_indkey(int64_t n, const mc_indel_rec* recs, const uint8_t* seq, uint64_t* keys, uint32_t* idx, mc_stream_t s)
{ if (n > 0) { mc_indkey_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(n, recs, seq, keys, idx); g_launches++; } }
static void launch_indlen(int64_t n, const mc_indel_rec* recs, const uint32_t* idx, uint32_t* len, mc_stream_t s)
{ if (n > 0) { mc_indlen_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(n, recs, idx, len); g_launches++; } }
static void launch_indgather(int64_t n, const mc_indel_rec* recs, const uint8_t* seq, const uint32_t* idx, const int64_t* off, mc_indel_rec* out, uint8_t* out_seq, mc_stream_t s)
{ if (n > 0) { mc_indgather_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(n, recs, seq, idx, off, out, out_seq); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_profhash_kernel(int64_t g0, int64_t n, const uint64_t* recs, mc_u64* acc)
{
	uint64_t sum = 0, x = 0;
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { const uint64_t h = profhash_of(g0 + i, recs, i); sum += h; x ^= h; }
	for (int o = 16; o; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); x ^= __shfl_xor_sync(0xffffffffu, x, o); }
	if ((threadIdx.x & 31) == 0) { atomicAdd(acc, (mc_u64)sum); atomicXor(acc + 1, (mc_u64)x); }
}
static void launch_profhash(int64_t g0, int64_t n, const uint64_t* recs, mc_u64* acc, mc_stream_t s)
{ if (n > 0) { int64_t b = (n + MC_BLOCK - 1) / MC_BLOCK; if (b > 148 * 8) b = 148 * 8; mc_profhash_kernel<<<(unsigned)b, MC_BLOCK, 0, s>>>(g0, n, recs, acc); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_gatecnt_kernel(const PipeArgs a, const ProfArgs q, int64_t n, uint64_t* list, mc_u64* bump)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) gatecnt_body(i, a, q, list, bump); }
static void launch_gatecnt(const PipeArgs& a, const ProfArgs& q, int64_t n, uint64_t* list, mc_u64* bump, mc_stream_t s)
{ if (n > 0) { mc_gatecnt_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, q, n, list, bump); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_gateadd_kernel(const PipeArgs a, int64_t n, const uint64_t* list)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) gateadd_body(i, a, list); }
static void launch_gateadd(const PipeArgs& a, int64_t n, const uint64_t* list, mc_stream_t s)
{ if (n > 0) { mc_gateadd_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, n, list); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_gatedense_fill_kernel(const PipeArgs a, const ProfArgs q, int64_t n, uint8_t* dense)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) gatedense_fill_body(i, a, q, dense); }
static void launch_gatedense_fill(const PipeArgs& a, const ProfArgs& q, int64_t n, uint8_t* dense, mc_stream_t s)
{ if (n > 0) { mc_gatedense_fill_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, q, n, dense); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_gatedense_apply_kernel(const PipeArgs a, int64_t G, const uint8_t* all, size_t pitch, int r0, int r1)
{ int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (g < G) gatedense_apply_body(g, a, all, pitch, r0, r1); }
static void launch_gatedense_apply(const PipeArgs& a, int64_t G, const uint8_t* all, size_t pitch, int r0, int r1, mc_stream_t s)
{ if (G > 0 && r1 > r0) { mc_gatedense_apply_kernel<<<(unsigned)((G + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, G, all, pitch, r0, r1); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_fqcount_kernel(const FastqArgs q, int f, int64_t n)
{ int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (t < n) fqcount_body(t, f, q); }
__global__ void __launch_bounds__(MC_BLOCK) mc_fqlines_kernel(const FastqArgs q, int f, int64_t n)
{ int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (t < n) fqlines_body(t, f, q); }
__global__ void __launch_bounds__(MC_BLOCK) mc_fqread_kernel(const FastqArgs q, int64_t n)
{ int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (r < n) fqread_body(r, q); }
__global__ void __launch_bounds__(MC_BLOCK) mc_fqcopy_kernel(const FastqArgs q, int64_t n)
{ const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; if (w < n) fqcopy_body(w, threadIdx.x & 31, 32, q); }
static void launch_fqcount(const FastqArgs& q, int f, int64_t n, mc_stream_t s)
{ if (n > 0) { mc_fqcount_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(q, f, n); g_launches++; } }
static void launch_fqlines(const FastqArgs& q, int f, int64_t n, mc_stream_t s)
{ if (n > 0) { mc_fqlines_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(q, f, n); g_launches++; } }
static void launch_fqread(const FastqArgs& q, int64_t n, mc_stream_t s)
{ if (n > 0) { mc_fqread_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(q, n); g_launches++; } }
static void launch_fqcopy(const FastqArgs& q, int64_t n, mc_stream_t s)
{ if (n > 0) { mc_fqcopy_kernel<<<(unsigned)((n * 32 + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(q, n); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_bwtsearch_kernel(const SearchArgs a, int64_t n)
{ int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (q < n) bwtsearch_body(q, a); }
static void launch_bwtsearch(const SearchArgs& a, int64_t n, mc_stream_t s)
{ if (n > 0) { mc_bwtsearch_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, n); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_cbwt_build_kernel(int64_t n, const uint32_t* src, uint32_t* dst)
{ int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (b < n) mc_cbwt_build_body(b, src, dst); }
static void launch_cbwt_build(int64_t n, const uint32_t* src, uint32_t* dst, mc_stream_t s)
{ if (n > 0) { mc_cbwt_build_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(n, src, dst); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_sa_dense_kernel(const DevIndex ix, int64_t n, int shift, uint32_t* o32, uint64_t* o64)
{ int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (j < n) mc_sa_dense_body(j, ix, shift, o32, o64); }
static void launch_sa_dense(const DevIndex& ix, int64_t n, int shift, uint32_t* o32, uint64_t* o64, mc_stream_t s)
{ if (n > 0) { mc_sa_dense_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(ix, n, shift, o32, o64); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_ktab_build_kernel(const DevIndex ix, int k, uint32_t* o32, uint64_t* o64)
{ int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (m < (1ll << (2 * k))) mc_ktab_build_body(m, ix, k, o32, o64); }
static void launch_ktab_build(const DevIndex& ix, int k, uint32_t* o32, uint64_t* o64, mc_stream_t s)
{ const int64_t n = 1ll << (2 * k); mc_ktab_build_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(ix, k, o32, o64); g_launches++; }
__global__ void __launch_bounds__(MC_BLOCK) mc_profsum_kernel(const DevProfile p, int64_t G, int64_t nb, int64_t b0, int64_t b1, int64_t* sums)
{ int64_t b = b0 + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5); if (b < b1) profsum_body(b, threadIdx.x & 31, 32, p, G, nb, sums); }   // whole warps leave together
static void launch_profsum(const DevProfile& p, int64_t G, int64_t nb, int64_t b0, int64_t b1, int64_t* sums, mc_stream_t s)
{ if (b1 > b0) { mc_profsum_kernel<<<(unsigned)(((b1 - b0) * 32 + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(p, G, nb, b0, b1, sums); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_profpack_kernel(const DevIndex ix, const DevProfile p, int64_t nb, const int64_t* pre, int64_t b0, int64_t b1, int64_t beg, int64_t end, uint64_t* out)
{ int64_t b = b0 + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5); if (b < b1) profpack_body(b, threadIdx.x & 31, 32, ix, p, nb, pre, beg, end, out); }
static void launch_profpack(const DevIndex& ix, const DevProfile& p, int64_t nb, const int64_t* pre, int64_t b0, int64_t b1, int64_t beg, int64_t end, uint64_t* out, mc_stream_t s)
{ if (b1 > b0) { mc_profpack_kernel<<<(unsigned)(((b1 - b0) * 32 + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(ix, p, nb, pre, b0, b1, beg, end, out); g_launches++; } }
#endif

MC_LAUNCH1(expand)
MC_LAUNCH1(cluster)
MC_LAUNCH1(single)
MC_LAUNCH1(pair)
MC_LAUNCH1B(alnprep)
#ifdef MC_HOSTEMU
MC_LAUNCH1(dp)
#else
#include "mc_dp_warp.cuh"
#endif
MC_LAUNCH1(alnfin)
MC_LAUNCH1(pairstat)
MC_LAUNCH2B(profkey)
MC_LAUNCH2(gate)
MC_LAUNCH2(gateupd)
MC_LAUNCH2B(scatter)
#ifndef MC_HOSTEMU
// persistent warps over the pieces queued by mc_scatter_kernel
__global__ void __launch_bounds__(MC_BLOCK) mc_profpiece_kernel(const PipeArgs a, const ProfArgs q)
{
	// 8-lane tiles, one piece each (no communication between the lanes of a tile)
	const int64_t tile = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3, n_tiles = ((int64_t)gridDim.x * blockDim.x) >> 3;
	const int64_t n = (int64_t)*a.ptask_bump;
	int natom = 0;
	for (int64_t t = tile; t < n; t += n_tiles) natom += profpiece_body(t, threadIdx.x & 7, 8, a, q);
	__syncwarp();
	mc_stat_add(&a.st->profile_atomics, (uint32_t)natom);
}
static void launch_profpiece(const PipeArgs& a, const ProfArgs& q, int64_t max_tasks, mc_stream_t s)
{
	if (max_tasks <= 0) return;
	int64_t blocks = (max_tasks * 8 + MC_BLOCK - 1) / MC_BLOCK; if (blocks > 148 * 8) blocks = 148 * 8;
	mc_profpiece_kernel<<<(unsigned)blocks, MC_BLOCK, 0, s>>>(a, q); g_launches++;
}
#endif

// ---- exclusive scan uint32 -> int64 (out has n + 1 entries) -----------------------------------------
#ifdef MC_HOSTEMU
static void device_scan_u32(const uint32_t* in, int64_t* out, int64_t n, int64_t*, mc_stream_t)
{ int64_t s = 0; for (int64_t i = 0; i < n; i++) { out[i] = s; s += in[i]; } out[n] = s; }
static size_t device_scan_scratch_bytes(int64_t) { return 8; }
static void device_sort_u64(uint64_t* keys, uint64_t*, int64_t n, void*, size_t, mc_stream_t) { std::sort(keys, keys + n); }
static void device_sort_pairs(uint64_t* keys, uint64_t*, uint32_t* vals, uint32_t*, int64_t n, void*, size_t, mc_stream_t)
{
	std::vector<std::pair<uint64_t, uint32_t> > v((size_t)n);
	for (int64_t i = 0; i < n; i++) v[(size_t)i] = std::make_pair(keys[i], vals[i]);
	std::stable_sort(v.begin(), v.end(), [](const std::pair<uint64_t, uint32_t>& x, const std::pair<uint64_t, uint32_t>& y) { return x.first < y.first; });
	for (int64_t i = 0; i < n; i++) { keys[i] = v[(size_t)i].first; vals[i] = v[(size_t)i].second; }
}
static size_t device_sort_pairs_scratch_bytes(int64_t) { return 8; }
static size_t device_sort_scratch_bytes(int64_t) { return 8; }
#else
#include "mc_scan.cuh"
// exclusive sums uint32 -> int64 (out has n + 1 entries: out[n] = total), in-place exclusive sums of int64 (+ total) and
// in-place inclusive maxima of int64: one decoupled look-back kernel each (mc_scan.cuh)
static void device_scan_u32(const uint32_t* in, int64_t* out, int64_t n, int64_t* scratch, mc_stream_t s)
{ device_lookback_scan<uint32_t, ScanSum, false>(in, out, n, scratch, out + n, s); g_launches++; }
static void device_exscan_i64(int64_t* a, int64_t n, int64_t* total, void* scratch, mc_stream_t s)
{ device_lookback_scan<int64_t, ScanSum, false>(a, a, n, scratch, total, s); g_launches++; }
// The variant-scan walkers (one thread per block of 100 columns) read the packed profile through shared memory: the thread
// block (128 walkers = 12800 columns) stages MC_VC_STAGE columns of every walker at a time with coalesced 16-byte loads
// (runs of 320 contiguous bytes), so that every 32-byte sector of the 16 B/column image is fetched from DRAM exactly once
// and whole; a row of 20 records is padded to 21 so that the walkers' own 16-byte reads spread over all banks.
#define MC_VC_THREADS 128
struct VcFetchShared {
	const uint4* tile; int64_t col0, n_cols;     // packed tile, tile-relative column of walker 0 of this thread block, columns in the tile
	uint4* sh;                                   // [MC_VC_THREADS][MC_VC_STAGE + 1]
	__device__ __forceinline__ void sync(int j)
	{
		if (j % MC_VC_STAGE) return;
		__syncthreads();                         // the previous stage has been consumed
		for (int e = threadIdx.x; e < MC_VC_THREADS * MC_VC_STAGE; e += MC_VC_THREADS)
		{
			const int seg = e / MC_VC_STAGE, within = e % MC_VC_STAGE;
			const int64_t col = col0 + (int64_t)seg * MC_VC_BLOCK + j + within;
			uint4 v = make_uint4(0, 0, 0, 0);
			if (col < n_cols) v = __ldg(tile + col);
			sh[seg * (MC_VC_STAGE + 1) + within] = v;
		}
		__syncthreads();
	}
	__device__ __forceinline__ void get(int j, uint64_t& w0, uint64_t& w1) const
	{
		const uint4 v = sh[threadIdx.x * (MC_VC_STAGE + 1) + j % MC_VC_STAGE];
		w0 = (uint64_t)v.y << 32 | v.x; w1 = (uint64_t)v.w << 32 | v.z;
	}
};
__global__ void __launch_bounds__(MC_VC_THREADS) mc_vcdepth_kernel(const VcArgs a, int64_t b0, int64_t b1)
{
	__shared__ uint4 sh[MC_VC_THREADS * (MC_VC_STAGE + 1)];
	const int64_t bb = b0 + blockIdx.x * (int64_t)MC_VC_THREADS, b = bb + threadIdx.x;
	VcFetchShared f; f.tile = (const uint4*)a.recs; f.col0 = bb * MC_VC_BLOCK - a.tile_beg; f.n_cols = a.tile_end - a.tile_beg; f.sh = sh;
	vcdepth_walk(b, b < b1, a, f);
}
static void launch_vcdepth(const VcArgs& a, int64_t b0, int64_t b1, mc_stream_t s)
{ if (b1 > b0) { mc_vcdepth_kernel<<<(unsigned)((b1 - b0 + MC_VC_THREADS - 1) / MC_VC_THREADS), MC_VC_THREADS, 0, s>>>(a, b0, b1); g_launches++; } }
__global__ void __launch_bounds__(MC_VC_THREADS) mc_vcscan_kernel(const VcArgs a, int64_t b0, int64_t b1, bool emit)
{
	__shared__ uint4 sh[MC_VC_THREADS * (MC_VC_STAGE + 1)];
	const int64_t bb = b0 + blockIdx.x * (int64_t)MC_VC_THREADS, b = bb + threadIdx.x;
	VcFetchShared f; f.tile = (const uint4*)a.recs; f.col0 = bb * MC_VC_BLOCK - a.tile_beg; f.n_cols = a.tile_end - a.tile_beg; f.sh = sh;
	vcscan_walk(b, b < b1, a, emit, f);
}
static void launch_vcscan(const VcArgs& a, int64_t b0, int64_t b1, bool emit, mc_stream_t s)
{ if (b1 > b0) { mc_vcscan_kernel<<<(unsigned)((b1 - b0 + MC_VC_THREADS - 1) / MC_VC_THREADS), MC_VC_THREADS, 0, s>>>(a, b0, b1, emit); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_vckey_kernel(int64_t n, const mc_variant_rec* recs, uint64_t* keys, uint32_t* idx, mc_u64* n_valid)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) vckey_body(i, recs, keys, idx, n_valid); }
static void launch_vckey(int64_t n, const mc_variant_rec* recs, uint64_t* keys, uint32_t* idx, mc_u64* n_valid, mc_stream_t s)
{ if (n > 0) { mc_vckey_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(n, recs, keys, idx, n_valid); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_vcgather_kernel(int64_t n, const mc_variant_rec* in, const uint32_t* idx, mc_variant_rec* out)
{ int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (j < n) vcgather_body(j, in, idx, out); }
static void launch_vcgather(int64_t n, const mc_variant_rec* in, const uint32_t* idx, mc_variant_rec* out, mc_stream_t s)
{ if (n > 0) { mc_vcgather_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(n, in, idx, out); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_samrec_kernel(const SamArgs a, int64_t n, bool emit)
{ int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (r < n) samrec_body(r, a, emit); }
static void launch_samrec(const SamArgs& a, int64_t n, bool emit, mc_stream_t s)
{ if (n > 0) { mc_samrec_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(a, n, emit); g_launches++; } }
__global__ void __launch_bounds__(MC_BLOCK) mc_samtext_kernel(const SamTextArgs t, int64_t n, bool emit)
{ int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (r < n) samtext_body(r, t, emit); }
static void launch_samtext(const SamTextArgs& t, int64_t n, bool emit, mc_stream_t s)
{ if (n > 0) { mc_samtext_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(t, n, emit); g_launches++; } }
static void device_incmax_i64(int64_t* a, int64_t n, void* scratch, mc_stream_t s)
{ if (n > 0) { device_lookback_scan<int64_t, ScanMax, true>(a, a, n, scratch, nullptr, s); g_launches++; } }
#include <cub/device/device_radix_sort.cuh>
static size_t device_sort_scratch_bytes(int64_t n)
{
	size_t b = 0; cub::DeviceRadixSort::SortKeys(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, n);
	return b + 256;
}
static size_t device_sort_pairs_scratch_bytes(int64_t n)
{
	size_t b = 0; cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
	return b + 256;
}
// sorts (key, value) pairs in place (the tmp arrays must hold n entries each); stable
static void device_sort_pairs(uint64_t* keys, uint64_t* ktmp, uint32_t* vals, uint32_t* vtmp, int64_t n, void* scratch, size_t scratch_bytes, mc_stream_t s)
{
	if (n <= 1) return;
	cub::DeviceRadixSort::SortPairs(scratch, scratch_bytes, (const uint64_t*)keys, ktmp, (const uint32_t*)vals, vtmp, n, 0, 64, s);
	cudaMemcpyAsync(keys, ktmp, (size_t)n * 8, cudaMemcpyDeviceToDevice, s); cudaMemcpyAsync(vals, vtmp, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
	g_launches += 4;
}
// sorts keys in place (tmp must hold n keys)
static void device_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, void* scratch, size_t scratch_bytes, mc_stream_t s)
{
	if (n <= 1) return;
	cub::DeviceRadixSort::SortKeys(scratch, scratch_bytes, (const uint64_t*)keys, tmp, n, 0, 64, s);
	cudaMemcpyAsync(keys, tmp, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
	g_launches += 4;
}
#endif

#endif
