// BWT and sampled suffix array of the index text on the GPU ("index construction").
//
// The CPU builder (index.cpp) sorts the suffixes of text = forward + reverse complement (2 bits per symbol, 32 symbols per
// 64-bit word, two zero words of padding; a suffix that runs out is smaller than one that goes on).  The suffix array is
// unique, so any correct sort yields the same .bwt / .sa files.  This one never materialises the suffix array: a GRCh38-sized
// text has 6.2 G suffixes (50 GB as 64-bit values), while the index only needs the BWT symbol of every row (2 bits) and the
// suffix of every 32nd row.  The rows are produced in CHUNKS of consecutive rows:
//   1. histogram of the first 12 symbols of every suffix (16 M bins) -> the rows of every 12-mer bin, chunk = a run of bins
//      holding at most `chunk_max` suffixes;
//   2. per chunk: select the suffixes whose 12-mer falls into the chunk (one pass over the packed text), key = the first 29
//      symbols (58 bits) above min(29, symbols left) so that a suffix that ends inside the window sorts before the one whose
//      next symbols equal the zero padding; one CUB radix sort of (key, position);
//   3. refinement rounds on what is still tied (repeats): the tied suffixes are compacted - their slots in the chunk stay
//      put, ascending -, keyed by (id of their group << 32 | next 14 symbols << 4 | min(14, symbols left)) and sorted again;
//      a suffix alone in its group is final.  Unique random text is done after step 2, a 3 kb exact repeat after ~200 cheap
//      rounds over a few thousand suffixes;
//   4. a final suffix at row r emits text[pos - 1] into the row-symbol array, its position into the sample array when
//      r % 32 == 0, and r as `primary` when pos == 0.
// Works for texts of up to 2^34 symbols (positions travel as 64-bit values); memory: ~1.6 bytes per text symbol plus 64 bytes
// per suffix of a chunk.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include "../../include/mapcaller_b200.h"

void mc_set_error(const char* fmt, ...);

namespace {

constexpr int kBinSyms = 12;                       // symbols of the chunking histogram
constexpr int64_t kBins = 1ll << (2 * kBinSyms);
constexpr int kSyms0 = 29, kSymsR = 14;            // symbols consumed by the first sort / by a refinement round

__device__ __forceinline__ uint64_t text_window(const uint64_t* __restrict__ w, int64_t i)
{
	const int64_t k = i >> 5; const int s = (int)(i & 31) << 1;
	return s ? (w[k] << s) | (w[k + 1] >> (64 - s)) : w[k];
}
__device__ __forceinline__ int text_base(const uint64_t* __restrict__ w, int64_t i) { return (int)(w[i >> 5] >> ((~i & 31) << 1)) & 3; }

__global__ void __launch_bounds__(256) bwt_hist_kernel(const uint64_t* __restrict__ w, int64_t n, unsigned long long* hist)
{
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i < n) atomicAdd(hist + (text_window(w, i) >> (64 - 2 * kBinSyms)), 1ull);
}

// suffixes whose first 12 symbols fall into [klo, khi): (key, position) appended in any order (block-aggregated cursor)
__global__ void __launch_bounds__(256) bwt_select_kernel(const uint64_t* __restrict__ w, int64_t n, uint32_t klo, uint32_t khi,
                                                         unsigned long long* cursor, uint64_t* keys, uint64_t* vals)
{
	__shared__ uint32_t s_cnt; __shared__ unsigned long long s_base;
	if (threadIdx.x == 0) s_cnt = 0;
	__syncthreads();
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	uint64_t win = 0; bool in = false;
	if (i < n) { win = text_window(w, i); const uint32_t b = (uint32_t)(win >> (64 - 2 * kBinSyms)); in = b >= klo && b < khi; }
	const uint32_t m = __ballot_sync(0xFFFFFFFFu, in), lane = threadIdx.x & 31;
	uint32_t wbase = 0;
	if (lane == 0 && m) wbase = atomicAdd(&s_cnt, (uint32_t)__popc(m));
	wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
	__syncthreads();
	if (threadIdx.x == 0 && s_cnt) s_base = atomicAdd(cursor, (unsigned long long)s_cnt);
	__syncthreads();
	if (in)
	{
		const uint64_t j = s_base + wbase + (uint32_t)__popc(m & ((1u << lane) - 1u));
		const int64_t left = n - i;
		keys[j] = (win & ~0x3Full) | (uint64_t)(left < kSyms0 ? left : kSyms0);
		vals[j] = (uint64_t)i;
	}
}

struct Emit {
	const uint64_t* w; int64_t n;
	uint8_t* rowsym;                 // symbol of sorted index S (row S + 1)
	uint64_t* samples;               // samples[j] = suffix of row 32 j
	unsigned long long* primary;     // row of the suffix at position 0
	int64_t row0;                    // sorted index of the chunk's slot 0
};
__device__ __forceinline__ void emit_final(const Emit& e, uint32_t slot, uint64_t pos)
{
	const int64_t S = e.row0 + slot, row = S + 1;
	if (pos == 0) { *e.primary = (unsigned long long)row; e.rowsym[S] = 0; }
	else e.rowsym[S] = (uint8_t)text_base(e.w, (int64_t)pos - 1);
	if ((row & 31) == 0) e.samples[row >> 5] = pos;
}

// after a sort: element c sits at slot (slot ? slot[c] : c); alone in its key run -> final, else flagged for the next round
__global__ void __launch_bounds__(256) bwt_resolve_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, const uint32_t* __restrict__ slot,
                                                          int64_t U, Emit e, uint32_t* tied, uint32_t* head)
{
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c >= U) return;
	const uint64_t k = keys[c];
	const bool h = c == 0 || keys[c - 1] != k, t = c + 1 == U || keys[c + 1] != k;
	if (h && t) { emit_final(e, slot ? slot[c] : (uint32_t)c, vals[c]); tied[c] = 0; head[c] = 0; }
	else { tied[c] = 1; head[c] = h ? 1u : 0u; }
}
// compaction of the tied elements: position, slot, and the compacted index of a group's head (0 elsewhere; max-scanned next)
__global__ void __launch_bounds__(256) bwt_compact_kernel(const uint64_t* __restrict__ vals, const uint32_t* __restrict__ slot, const uint32_t* __restrict__ tied,
                                                          const uint32_t* __restrict__ head, const uint32_t* __restrict__ off, int64_t U,
                                                          uint64_t* vals2, uint32_t* slot2, uint32_t* gid2)
{
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c >= U || !tied[c]) return;
	const uint32_t j = off[c];
	vals2[j] = vals[c]; slot2[j] = slot ? slot[c] : (uint32_t)c; gid2[j] = head[c] ? j : 0u;
}
__global__ void __launch_bounds__(256) bwt_rekey_kernel(const uint64_t* __restrict__ w, int64_t n, const uint64_t* __restrict__ vals, const uint32_t* __restrict__ gid,
                                                        int64_t U, int64_t depth, uint64_t* keys)
{
	const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (c >= U) return;
	const int64_t p = (int64_t)vals[c] + depth;
	uint64_t sym = 0, left = 0;
	if (p < n) { sym = text_window(w, p) >> (64 - 2 * kSymsR); left = (uint64_t)(n - p < kSymsR ? n - p : kSymsR); }
	keys[c] = (uint64_t)gid[c] << 32 | sym << 4 | left;
}
__global__ void __launch_bounds__(256) bwt_pack_kernel(const uint8_t* __restrict__ rowsym, int64_t n, uint64_t* out)
{
	const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (32 * k >= n) return;
	uint64_t word = 0;
	const int64_t p1 = 32 * k + 32 < n ? 32 * k + 32 : n;
	for (int64_t p = 32 * k; p < p1; p++) word |= (uint64_t)(rowsym[p] & 3) << ((~p & 31) << 1);
	out[k] = word;
}

struct MaxOp { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

struct Bufs {
	void* p[32]; int n = 0; size_t total = 0;
	void* get(size_t bytes) { void* q = nullptr; if (cudaMalloc(&q, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; } p[n++] = q; total += bytes; return q; }
	~Bufs() { for (int i = 0; i < n; i++) cudaFree(p[i]); }
};
inline unsigned grid_of(int64_t items) { return (unsigned)((items + 255) / 256); }

} // namespace

// words: the packed text of index.cpp (n_words 64-bit words, padding included).  Outputs on the host:
//   rowsym  (n + 31) / 32 words: BWT symbol of row S + 1 at symbol S, packed like the text (0 where the row is `primary`)
//   primary the row of the suffix that starts at position 0
//   samples n / 32 + 1 entries: samples[j] = suffix of row 32 j (samples[0] = -1: row 0 is the empty suffix)
extern "C" int mc_gpu_bwt_build(const uint64_t* words, size_t n_words, int64_t n, int device, uint64_t* rowsym_out, uint64_t* primary_out, uint64_t* samples_out)
{
	if (n <= 0 || n >= (1ll << 34)) { mc_set_error("mc_index_build_gpu: the text must have fewer than 2^34 symbols"); return MC_ERR_ARG; }
	if (cudaSetDevice(device) != cudaSuccess) { mc_set_error("mc_index_build_gpu: no CUDA device %d", device); return MC_ERR_CUDA; }
	const bool verbose = getenv("MC_DEBUG") != nullptr;
	Bufs b;
	const int64_t n_samples = n / 32 + 1, n_pack = (n + 31) / 32;
	uint64_t* d_w = (uint64_t*)b.get(n_words * 8);
	uint8_t* d_rowsym = (uint8_t*)b.get((size_t)n);
	uint64_t* d_samples = (uint64_t*)b.get((size_t)n_samples * 8);
	uint64_t* d_pack = (uint64_t*)b.get((size_t)n_pack * 8);
	unsigned long long* d_hist = (unsigned long long*)b.get((size_t)kBins * 8);
	unsigned long long* d_small = (unsigned long long*)b.get(64);     // [0] select cursor, [1] primary
	if (!d_w || !d_rowsym || !d_samples || !d_pack || !d_hist || !d_small) { mc_set_error("mc_index_build_gpu: out of device memory"); return MC_ERR_CUDA; }
	size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
	int64_t chunk_max = (int64_t)((free_b > (3ull << 30) ? free_b - (2ull << 30) : free_b / 2) / 72);
	chunk_max = std::min<int64_t>(chunk_max, (1ll << 30) - 1);
	if (const char* e = getenv("MC_INDEX_CHUNK")) chunk_max = std::max<int64_t>(1, std::min<int64_t>(chunk_max, atoll(e)));   // tests: force several chunks
	const int64_t cap = std::min<int64_t>(chunk_max, n);
	uint64_t* keys[2] = {(uint64_t*)b.get((size_t)cap * 8), (uint64_t*)b.get((size_t)cap * 8)};
	uint64_t* vals[2] = {(uint64_t*)b.get((size_t)cap * 8), (uint64_t*)b.get((size_t)cap * 8)};
	uint32_t* slot[2] = {(uint32_t*)b.get((size_t)cap * 4), (uint32_t*)b.get((size_t)cap * 4)};
	uint32_t* tied = (uint32_t*)b.get((size_t)cap * 4);
	uint32_t* head = (uint32_t*)b.get((size_t)cap * 4);
	uint32_t* off = (uint32_t*)b.get((size_t)cap * 4);
	uint32_t* gid[2] = {(uint32_t*)b.get((size_t)cap * 4), (uint32_t*)b.get((size_t)cap * 4)};
	size_t t_sort = 0, t_scan = 0, t_max = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, t_sort, keys[0], keys[1], vals[0], vals[1], (int)cap, 0, 64);
	cub::DeviceScan::ExclusiveSum(nullptr, t_scan, tied, off, (int)cap);
	cub::DeviceScan::InclusiveScan(nullptr, t_max, gid[0], gid[1], MaxOp(), (int)cap);
	const size_t t_bytes = std::max(t_sort, std::max(t_scan, t_max));
	void* d_tmp = b.get(t_bytes);
	if (!keys[0] || !keys[1] || !vals[0] || !vals[1] || !slot[0] || !slot[1] || !tied || !head || !off || !gid[0] || !gid[1] || !d_tmp)
	{ mc_set_error("mc_index_build_gpu: out of device memory (%.1f GB wanted)", (double)b.total / 1e9); return MC_ERR_CUDA; }

	cudaMemcpy(d_w, words, n_words * 8, cudaMemcpyHostToDevice);
	cudaMemset(d_hist, 0, (size_t)kBins * 8);
	cudaMemset(d_small, 0, 64);
	cudaMemset(d_samples, 0xFF, 8);                                    // samples[0] = -1
	bwt_hist_kernel<<<grid_of(n), 256>>>(d_w, n, d_hist);
	std::vector<unsigned long long> hist((size_t)kBins);
	if (cudaMemcpy(hist.data(), d_hist, (size_t)kBins * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { mc_set_error("mc_index_build_gpu: %s", cudaGetErrorString(cudaGetLastError())); return MC_ERR_CUDA; }

	cudaEvent_t ev[4]; for (auto& x : ev) cudaEventCreate(&x);
	float ms_select = 0, ms_sort0 = 0, ms_refine = 0;
	Emit em; em.w = d_w; em.n = n; em.rowsym = d_rowsym; em.samples = d_samples; em.primary = d_small + 1;
	int64_t row0 = 0, bin = 0; int n_chunks = 0; long n_rounds = 0;
	while (bin < kBins)
	{
		int64_t m = 0, b1 = bin;
		while (b1 < kBins && m + (int64_t)hist[b1] <= cap) m += (int64_t)hist[b1++];
		if (b1 == bin) { mc_set_error("mc_index_build_gpu: %llu suffixes share their first %d symbols, more than one chunk holds (%lld)", hist[bin], kBinSyms, (long long)cap); return MC_ERR_ARG; }
		if (m > 0)
		{
			n_chunks++;
			em.row0 = row0;
			cudaMemsetAsync(d_small, 0, 8);
			cudaEventRecord(ev[0]);
			bwt_select_kernel<<<grid_of(n), 256>>>(d_w, n, (uint32_t)bin, (uint32_t)b1, d_small, keys[0], vals[0]);
			cudaEventRecord(ev[1]);
			size_t tb = t_bytes;
			cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys[0], keys[1], vals[0], vals[1], (int)m, 0, 64);
			cudaEventRecord(ev[2]);
			// the sorted pairs are in [1]; slots are implicit in the first round
			int64_t U = m, depth = kSyms0; int cur = 1; const uint32_t* cur_slot = nullptr; int sl = 0;
			for (;;)
			{
				n_rounds++;
				bwt_resolve_kernel<<<grid_of(U), 256>>>(keys[cur], vals[cur], cur_slot, U, em, tied, head);
				tb = t_bytes; cub::DeviceScan::ExclusiveSum(d_tmp, tb, tied, off, (int)U);
				uint32_t last[2] = {0, 0};
				if (cudaMemcpy(&last[0], off + (U - 1), 4, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(&last[1], tied + (U - 1), 4, cudaMemcpyDeviceToHost) != cudaSuccess)
				{ mc_set_error("mc_index_build_gpu: suffix sort failed (%s)", cudaGetErrorString(cudaGetLastError())); return MC_ERR_CUDA; }
				const int64_t U2 = (int64_t)last[0] + last[1];
				if (U2 == 0) break;
				const int nxt = cur ^ 1;
				bwt_compact_kernel<<<grid_of(U), 256>>>(vals[cur], cur_slot, tied, head, off, U, vals[nxt], slot[sl], gid[0]);
				tb = t_bytes; cub::DeviceScan::InclusiveScan(d_tmp, tb, gid[0], gid[1], MaxOp(), (int)U2);
				bwt_rekey_kernel<<<grid_of(U2), 256>>>(d_w, n, vals[nxt], gid[1], U2, depth, keys[nxt]);
				int bits_u = 1; while ((1ll << bits_u) < U2) bits_u++;
				tb = t_bytes; cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys[nxt], keys[cur], vals[nxt], vals[cur], (int)U2, 0, 32 + bits_u);
				cur_slot = slot[sl]; sl ^= 1; U = U2; depth += kSymsR;
			}
			cudaEventRecord(ev[3]); cudaEventSynchronize(ev[3]);
			float x = 0; cudaEventElapsedTime(&x, ev[0], ev[1]); ms_select += x; cudaEventElapsedTime(&x, ev[1], ev[2]); ms_sort0 += x; cudaEventElapsedTime(&x, ev[2], ev[3]); ms_refine += x;
		}
		row0 += m; bin = b1;
	}
	bwt_pack_kernel<<<grid_of(n_pack), 256>>>(d_rowsym, n, d_pack);
	unsigned long long prim = 0;
	cudaMemcpy(rowsym_out, d_pack, (size_t)n_pack * 8, cudaMemcpyDeviceToHost);
	cudaMemcpy(samples_out, d_samples, (size_t)n_samples * 8, cudaMemcpyDeviceToHost);
	cudaMemcpy(&prim, d_small + 1, 8, cudaMemcpyDeviceToHost);
	const cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess || cudaGetLastError() != cudaSuccess || row0 != n || prim == 0)
	{ mc_set_error("mc_index_build_gpu: suffix sort failed (%s, %lld of %lld rows)", cudaGetErrorString(e), (long long)row0, (long long)n); return MC_ERR_CUDA; }
	*primary_out = (uint64_t)prim;
	for (auto& x : ev) cudaEventDestroy(x);
	if (verbose) fprintf(stderr, "[mc] index build on the GPU: %lld suffixes, %d chunk(s) of <= %lld, %ld sort rounds, %.1f GB of HBM; select %.0f ms, first sort %.0f ms, refinement %.0f ms\n",
	                     (long long)n, n_chunks, (long long)cap, n_rounds, (double)b.total / 1e9, ms_select, ms_sort0, ms_refine);
	return MC_OK;
}
