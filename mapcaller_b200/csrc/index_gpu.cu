// Suffix array of the index text on the GPU ("index construction"): prefix doubling with CUB radix sorts.
//
// The CPU builder (index.cpp) sorts the suffixes of text = forward + reverse complement (2 bits per symbol, 32 symbols per
// 64-bit word, two zero words of padding; a suffix that runs out is smaller than one that goes on).  The suffix array is
// unique, so any correct sort yields the same .bwt / .sa files; this one keeps the text, the keys and the ranks in HBM:
//   round 0: key = first 16 symbols (32 bits) and, below them, min(16, symbols left) so that a suffix that ends inside the
//            window sorts before the one whose next symbols equal the zero padding;
//   round k: key = rank[i] (by the first h symbols) : rank[i + h] (0 past the end), h = 16 * 2^(k-1);
// until every rank is unique.  36 bytes of HBM per text symbol; texts of 2^32 symbols and more stay with the CPU sort.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include "../../include/mapcaller_b200.h"

void mc_set_error(const char* fmt, ...);

namespace {

__global__ void __launch_bounds__(256) sufsort_init(const uint64_t* w, int64_t n, uint64_t* keys, uint32_t* vals)
{
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int64_t k = i >> 5; const int s = (int)(i & 31) << 1;
	const uint64_t win = s ? (w[k] << s) | (w[k + 1] >> (64 - s)) : w[k];
	const int64_t left = n - i;
	keys[i] = ((win >> 32) << 5) | (uint64_t)(left < 16 ? left : 16);
	vals[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) sufsort_flags(const uint64_t* keys, int64_t n, uint32_t* flags)
{
	const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (j < n) flags[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) sufsort_scatter(const uint32_t* sa, const uint32_t* rank_sorted, int64_t n, uint32_t* rank)
{
	const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (j < n) rank[sa[j]] = rank_sorted[j];
}
__global__ void __launch_bounds__(256) sufsort_keys(const uint32_t* rank, int64_t n, int64_t h, uint64_t* keys, uint32_t* vals)
{
	const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	keys[i] = ((uint64_t)rank[i] << 32) | (uint64_t)(i + h < n ? rank[i + h] : 0u);
	vals[i] = (uint32_t)i;
}

struct Bufs {
	void* p[16]; int n = 0;
	void* get(size_t bytes) { void* q = nullptr; if (cudaMalloc(&q, bytes ? bytes : 1) != cudaSuccess) return nullptr; p[n++] = q; return q; }
	~Bufs() { for (int i = 0; i < n; i++) cudaFree(p[i]); }
};

} // namespace

// words: the packed text of index.cpp (n_words 64-bit words, padding included); sa_out: n entries on the host
extern "C" int mc_gpu_suffix_sort(const uint64_t* words, size_t n_words, int64_t n, int device, uint32_t* sa_out)
{
	if (n <= 0 || n >= (1ll << 32) - 2) { mc_set_error("mc_index_build_gpu: the text must have fewer than 2^32 symbols"); return MC_ERR_ARG; }
	if (cudaSetDevice(device) != cudaSuccess) { mc_set_error("mc_index_build_gpu: no CUDA device %d", device); return MC_ERR_CUDA; }
	Bufs b;
	uint64_t* d_w = (uint64_t*)b.get(n_words * 8);
	uint64_t* keys[2] = {(uint64_t*)b.get((size_t)n * 8), (uint64_t*)b.get((size_t)n * 8)};
	uint32_t* vals[2] = {(uint32_t*)b.get((size_t)n * 4), (uint32_t*)b.get((size_t)n * 4)};
	uint32_t* rank = (uint32_t*)b.get((size_t)n * 4);
	uint32_t* flags = (uint32_t*)b.get((size_t)n * 4);
	uint32_t* rsorted = (uint32_t*)b.get((size_t)n * 4);
	size_t t_sort = 0, t_scan = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, t_sort, keys[0], keys[1], vals[0], vals[1], n, 0, 64);
	cub::DeviceScan::InclusiveSum(nullptr, t_scan, flags, rsorted, n);
	const size_t t_bytes = t_sort > t_scan ? t_sort : t_scan;
	void* d_tmp = b.get(t_bytes);
	if (!d_w || !keys[0] || !keys[1] || !vals[0] || !vals[1] || !rank || !flags || !rsorted || !d_tmp) { cudaGetLastError(); mc_set_error("mc_index_build_gpu: out of device memory (36 bytes per text symbol)"); return MC_ERR_CUDA; }
	const unsigned grid = (unsigned)((n + 255) / 256);
	int bits_n = 1; while ((1ll << bits_n) <= n) bits_n++;                    // ranks are 1..n
	cudaMemcpy(d_w, words, n_words * 8, cudaMemcpyHostToDevice);
	sufsort_init<<<grid, 256>>>(d_w, n, keys[0], vals[0]);
	int end_bit = 37;
	int64_t h = 16;
	for (int round = 0; round < 48; round++)
	{
		size_t tb = t_bytes;
		cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys[0], keys[1], vals[0], vals[1], n, 0, end_bit);
		sufsort_flags<<<grid, 256>>>(keys[1], n, flags);
		tb = t_bytes;
		cub::DeviceScan::InclusiveSum(d_tmp, tb, flags, rsorted, n);
		uint32_t last = 0;
		if (cudaMemcpy(&last, rsorted + (n - 1), 4, cudaMemcpyDeviceToHost) != cudaSuccess) break;
		if ((int64_t)last == n)
		{
			if (cudaMemcpy(sa_out, vals[1], (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) break;
			return MC_OK;
		}
		sufsort_scatter<<<grid, 256>>>(vals[1], rsorted, n, rank);
		sufsort_keys<<<grid, 256>>>(rank, n, h, keys[0], vals[0]);
		end_bit = 32 + bits_n;
		h *= 2;
	}
	const cudaError_t e = cudaGetLastError();
	mc_set_error("mc_index_build_gpu: suffix sort failed (%s)", e == cudaSuccess ? "no convergence" : cudaGetErrorString(e));
	return MC_ERR_CUDA;
}
