// Device abstraction + shared POD types of the mapping pipeline.
//
// The pipeline stages are written as per-work-item bodies (mc_stages.h) that nvcc compiles into
// sm_100a kernels.  Defining MC_HOSTEMU compiles the same bodies as plain C++ loops: that build
// (tools/hostemu) is a DEVELOPER HARNESS for debugging stage logic against the reference on a box
// without a GPU.  It is never linked into libmapcaller_b200.so, never loaded by the package, the
// tests or bench.py, and is not a fallback.
#ifndef MC_DEVICE_H
#define MC_DEVICE_H

#include <stdint.h>
#include <string.h>

#include "../../include/mapcaller_b200.h"

#ifdef MC_HOSTEMU
#include <stdio.h>
#include <stdlib.h>
#define MC_HD inline
#define MC_HOST_HD inline
#define MC_DEV_ONLY 0
struct mc_u32x4 { uint32_t x, y, z, w; };
static inline mc_u32x4 mc_ldg128(const void* p) { mc_u32x4 v; memcpy(&v, p, 16); return v; }
template <class T> static inline T mc_ldg(const T* p) { return *p; }
static inline mc_u32x4 mc_gather128(const void* p) { return mc_ldg128(p); }
static inline uint32_t mc_gather32(const uint32_t* p) { return *p; }
static inline uint64_t mc_gather64(const uint64_t* p) { return *p; }
template <class T> static inline T mc_atomic_add(T* p, T v) { T o = *p; *p = (T)(o + v); return o; }
template <class T> static inline void mc_atomic_or(T* p, T v) { *p = (T)(*p | v); }
static inline int mc_atomic_exch(int* p, int v) { int o = *p; *p = v; return o; }
static inline int mc_popc(uint32_t x) { return __builtin_popcount(x); }
static inline int mc_ctz(uint32_t x) { return __builtin_ctz(x); }
#define MC_WARP_SYNC() do { } while (0)
static inline int64_t mc_bcast64(int64_t v) { return v; }
static inline int mc_warp_sum(int v) { return v; }
static inline int mc_warp_max(int v) { return v; }
static inline int mc_max3(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }
static inline int mc_warp_incl_scan(int v, int) { return v; }
static inline int mc_warp_last(int v) { return v; }
#else
#include <cuda_runtime.h>
#define MC_HD __device__ __forceinline__
#define MC_HOST_HD __host__ __device__ __forceinline__
#define MC_DEV_ONLY 1
typedef uint4 mc_u32x4;
static __device__ __forceinline__ mc_u32x4 mc_ldg128(const void* p) { return __ldg((const uint4*)p); }
template <class T> static __device__ __forceinline__ T mc_ldg(const T* p) { return __ldg(p); }
// Random gathers from tables far larger than L2 (FM-index blocks, suffix-array samples, k-mer table): by default an L2 miss
// fetches the whole 128-byte line from HBM; the .L2::64B prefetch-size qualifier halves that (tools/probes/gather_probe2.cu:
// 123 -> 63 bytes of DRAM traffic per random 32-byte gather, no other load flavour changes it).
static __device__ __forceinline__ mc_u32x4 mc_gather128(const void* p)
{
	mc_u32x4 v;
	asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}
static __device__ __forceinline__ uint32_t mc_gather32(const uint32_t* p) { uint32_t v; asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
static __device__ __forceinline__ uint64_t mc_gather64(const uint64_t* p) { uint64_t v; asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }
static __device__ __forceinline__ unsigned long long mc_atomic_add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
static __device__ __forceinline__ uint32_t mc_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
static __device__ __forceinline__ int mc_atomic_add(int* p, int v) { return atomicAdd(p, v); }
static __device__ __forceinline__ void mc_atomic_or(unsigned long long* p, unsigned long long v) { atomicOr(p, v); }
static __device__ __forceinline__ void mc_atomic_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
static __device__ __forceinline__ int mc_atomic_exch(int* p, int v) { return atomicExch(p, v); }
static __device__ __forceinline__ int mc_popc(uint32_t x) { return __popc(x); }
static __device__ __forceinline__ int mc_ctz(uint32_t x) { return __ffs((int)x) - 1; }
#define MC_WARP_SYNC() __syncwarp()
static __device__ __forceinline__ int64_t mc_bcast64(int64_t v) { return __shfl_sync(0xffffffffu, v, 0); }
static __device__ __forceinline__ int mc_warp_sum(int v) { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }
static __device__ __forceinline__ int mc_warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
static __device__ __forceinline__ int mc_max3(int a, int b, int c) { return __vimax3_s32(a, b, c); } // DPX
// inclusive prefix sum over the 32 lanes of a (full) warp; mc_warp_last = the value lane 31 holds
static __device__ __forceinline__ int mc_warp_incl_scan(int v, int lane)
{
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
	return v;
}
static __device__ __forceinline__ int mc_warp_last(int v) { return __shfl_sync(0xffffffffu, v, 31); }
#endif

typedef unsigned long long mc_u64; // atomicAdd-compatible 64-bit counter

// 8-lane tiles of a warp (one normal piece each, mc_stages_align.h): sync and sum inside the tile
#ifdef MC_HOSTEMU
#define MC_TILE8_SYNC() do { } while (0)
static inline int mc_tile8_sum(int v) { return v; }
#else
#define MC_TILE8_MASK (0xFFu << (threadIdx.x & 24))
#define MC_TILE8_SYNC() __syncwarp(MC_TILE8_MASK)
static __device__ __forceinline__ int mc_tile8_sum(int v)
{
	const unsigned m = MC_TILE8_MASK;
	v += __shfl_xor_sync(m, v, 4); v += __shfl_xor_sync(m, v, 2); v += __shfl_xor_sync(m, v, 1);
	return v;
}
#endif

// "group" = all threads that work on one item together; for the rescue windows that is a whole thread block
#ifdef MC_HOSTEMU
#define MC_GROUP_SYNC() do { } while (0)
static inline int mc_group_any(int v) { return v; }
static inline int64_t mc_group_bcast64(int64_t v, int) { return v; }
#else
#define MC_GROUP_SYNC() __syncthreads()
static __device__ __forceinline__ int mc_group_any(int v) { return __syncthreads_or(v); }
static __device__ __forceinline__ int64_t mc_group_bcast64(int64_t v, int lane)
{
	__shared__ long long slot;
	if (lane == 0) slot = v;
	__syncthreads();
	v = slot;
	__syncthreads();
	return v;
}
#endif

// Arena cursors are bumped by (nearly) every thread of a kernel.  Same-address atomics serialise in the L2 atomic unit
// (about one lane per clock), so the lanes that arrive together first scan their sizes inside the warp and issue ONE
// atomic for the group; each lane then owns [base + prefix, base + prefix + n).
#ifdef MC_HOSTEMU
static inline int64_t mc_bump_alloc(mc_u64* bump, uint32_t n) { mc_u64 o = *bump; *bump += n; return (int64_t)o; }
static inline void mc_bump_alloc2(mc_u64* b0, uint32_t n0, mc_u64* b1, uint32_t n1, int64_t* o0, int64_t* o1) { *o0 = mc_bump_alloc(b0, n0); *o1 = mc_bump_alloc(b1, n1); }
#else
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
static __device__ __forceinline__ int64_t mc_bump_alloc(mc_u64* bump, uint32_t n)
{
	namespace cg = cooperative_groups;
	cg::coalesced_group g = cg::coalesced_threads();
	const uint32_t pre = cg::exclusive_scan(g, n);
	const uint32_t total = g.shfl(pre + n, g.size() - 1);
	mc_u64 base = 0;
	if (g.thread_rank() == 0) base = atomicAdd(bump, (mc_u64)total);
	base = g.shfl(base, 0);
	return (int64_t)(base + pre);
}
// two cursors at once: both atomics are in flight together, one round trip instead of two
static __device__ __forceinline__ void mc_bump_alloc2(mc_u64* b0, uint32_t n0, mc_u64* b1, uint32_t n1, int64_t* o0, int64_t* o1)
{
	namespace cg = cooperative_groups;
	cg::coalesced_group g = cg::coalesced_threads();
	const uint32_t p0 = cg::exclusive_scan(g, n0), p1 = cg::exclusive_scan(g, n1);
	const uint32_t t0 = g.shfl(p0 + n0, g.size() - 1), t1 = g.shfl(p1 + n1, g.size() - 1);
	mc_u64 a0 = 0, a1 = 0;
	if (g.thread_rank() == 0) { a0 = atomicAdd(b0, (mc_u64)t0); a1 = atomicAdd(b1, (mc_u64)t1); }
	a0 = g.shfl(a0, 0); a1 = g.shfl(a1, 0);
	*o0 = (int64_t)(a0 + p0); *o1 = (int64_t)(a1 + p1);
}
#endif

// Statistics counters are hit by every thread of a kernel: summing inside the warp first leaves one atomic per warp
// instead of 32 same-address atomics (which serialise in the L2 atomic unit).
#ifdef MC_HOSTEMU
static inline void mc_stat_add(mc_u64* p, uint32_t v) { *p += v; }
#else
static __device__ __forceinline__ void mc_stat_add(mc_u64* p, uint32_t v)
{
	const unsigned m = __activemask();
	const unsigned sum = __reduce_add_sync(m, v);
	if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1) && sum) atomicAdd(p, (mc_u64)sum);
}
#endif

// The same for kernels with one thread per read, where (nearly) every thread of a block needs a range: the sizes are scanned
// across the BLOCK and one thread issues the atomic - returning atomics on one address complete at roughly one per 20 clocks
// on B200 (the SHFL that waits for the warp's atomic was 57 % of the stall samples of the scatter kernel), so the count of
// atomics is what matters.  EVERY thread of the block must call (n = 0 when it needs nothing).  Up to two cursors at once.
#ifdef MC_HOSTEMU
static inline void mc_block_bump2(mc_u64* b0, uint32_t n0, mc_u64* b1, uint32_t n1, int64_t* o0, int64_t* o1)
{ *o0 = mc_bump_alloc(b0, n0); *o1 = b1 ? mc_bump_alloc(b1, n1) : 0; }
#else
static __device__ __forceinline__ void mc_block_bump2(mc_u64* b0, uint32_t n0, mc_u64* b1, uint32_t n1, int64_t* o0, int64_t* o1)
{
	__shared__ uint32_t wsum0[32], wsum1[32];
	__shared__ unsigned long long base0, base1;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	uint32_t v0 = n0, v1 = n1;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const uint32_t t0 = __shfl_up_sync(0xffffffffu, v0, d), t1 = __shfl_up_sync(0xffffffffu, v1, d);
		if (lane >= d) { v0 += t0; v1 += t1; }
	}
	if (lane == 31) { wsum0[w] = v0; wsum1[w] = v1; }
	__syncthreads();
	if (w == 0)
	{
		uint32_t s0 = lane < nw ? wsum0[lane] : 0, s1 = lane < nw ? wsum1[lane] : 0;
		const uint32_t own0 = s0, own1 = s1;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const uint32_t t0 = __shfl_up_sync(0xffffffffu, s0, d), t1 = __shfl_up_sync(0xffffffffu, s1, d);
			if (lane >= d) { s0 += t0; s1 += t1; }
		}
		wsum0[lane] = s0 - own0; wsum1[lane] = s1 - own1;     // exclusive prefix of the warp totals
		if (lane == 31)
		{
			base0 = s0 ? atomicAdd(b0, (mc_u64)s0) : 0;
			base1 = (b1 && s1) ? atomicAdd(b1, (mc_u64)s1) : 0;
		}
	}
	__syncthreads();
	*o0 = (int64_t)(base0 + wsum0[w] + v0 - n0);
	*o1 = (int64_t)(base1 + wsum1[w] + v1 - n1);
	__syncthreads();                                       // the scratch may be reused by the next call
}
#endif
static
#ifndef MC_HOSTEMU
__device__ __forceinline__
#else
inline
#endif
int64_t mc_block_bump(mc_u64* bump, uint32_t n) { int64_t o0, o1; mc_block_bump2(bump, n, nullptr, 0, &o0, &o1); return o0; }

// ---- FM-index replica in HBM (layout of the reference's bwt_t, unchanged) ------------------------
struct DevIndex {
	const uint32_t* bwt;   // reference layout, 64-byte blocks: 4 x uint64 occ counts + 8 x uint32 packed symbols (128 rows)
	const uint32_t* cbwt;  // compact device layout (texts < 2^32 symbols), 32-byte blocks of 64 rows: 4 x uint32 counts + low-bit plane + high-bit plane; 0 = not built
	const uint64_t* sa;    // sampled suffix array, one entry per (1 << sa_shift) rows: the reference's (every 32nd row) or, for texts >= 2^32 symbols, the denser device copy
	const uint32_t* sa32;  // denser device copy for texts < 2^32 symbols (32-bit entries); 0 = use `sa`
	int32_t sa_shift;      // log2 of the sampling interval of whichever table is in use
	const uint32_t* ktab32; // k-mer start table of the seed search (mc_fmindex.h), 2 x uint32 per k-mer (compact layout); 0 = none
	const uint64_t* ktab64; // the same with 2 x uint64 per k-mer (texts >= 2^32 symbols)
	int32_t ktab_k;         // k (0 = no table)
	const uint8_t* pac;    // forward 2-bit text; revcomp half derived on the fly
	const int64_t* chrom_end; // sorted keys of PosChrIdMap (reference src/bwt_index.cpp:253-254)
	const int32_t* chrom_id;  // value of each key
	int32_t n_end;
	uint64_t primary, L2[5], seq_len;
	int64_t G, twoG;
};

struct Seed { uint64_t x0; int32_t read; int16_t rpos; int16_t len; };   // one recorded BWT_Search hit: rows [x0, x0+freq) of the seed's reverse complement
struct SPair { int64_t gpos; int32_t rpos; int32_t len; };               // simple pair (exact-match seed placed on the genome)
struct Cand { int32_t score; int32_t pbeg; int32_t pend; };              // cluster = slice of the read's sorted simple pairs
struct DpTask { int32_t frag; int32_t m; int32_t n; int32_t pad; int64_t ws_off; };

struct DevStats {
	mc_u64 seed_blocks, seed_locate_blocks, seed_sa_reads;   // written by the seed kernel only (kept when the later stages are repeated with larger arenas)
	mc_u64 locate_blocks, sa_reads, dp_cells, dp_tasks, profile_columns, profile_atomics;
	mc_u64 overflow;      // any arena ran out: the batch is re-run with larger arenas
	mc_u64 odd_merge;     // defensive counter: IdentifyNormalPairs ordering assumption violated
};

struct DevParams {
	int32_t paired, alg_ksw2, max_pos_diff, max_clip, max_dup, update_profile;
	float max_mismatch_rate;
};

// device-resident pile-up profile (DESIGN.md "data layout").  Everything a read adds over a RANGE of columns
// (strand coverage, multi-hit coverage, exact-match seeds that by construction carry the reference base) is kept as
// a difference array (+1 at the first column, -1 one past the last) and only summed up when the profile is read
// out; per-column counters exist only for bases that may differ from the reference.
struct DevProfile {
	uint32_t* base16;  // [G][2] words = 4 x uint16: A,C | G,T  (bases of gapped / mismatching pieces)
	int32_t* sdiff;    // [G+1][4] difference arrays of F1, R2, F2, R1
	int32_t* cdiff;    // [G+1] difference array: reads whose exact-match seed covers the column (counts the reference base)
	int32_t* mdiff;    // [G+1] difference array of multi_hit
	uint8_t* rcount;   // [G] readCount gate (<= iMaxDuplicate)
};

#define MC_PROF_BLOCK 1024   // columns per read-out block

#endif
