// Single-pass device-wide scans with decoupled look-back (CUDA only; the developer harness uses serial loops, mc_launch.h).
//
// Every scan of the pipeline - seed-slot / FASTQ-line / CIGAR offsets (uint32 -> int64 exclusive sums), the block prefixes
// of the profile's difference arrays (int64 exclusive sums, 1 per 1024 columns) and the gap / duplication carriers of the
// variant-calling scan (int64 inclusive maxima, 1 per 100 columns) - runs as ONE kernel whose tiles of 2048 items are
// handed out by an atomic ticket (so a tile's predecessors are always already running), publish their aggregate, and
// resolve their exclusive prefix by looking back over the published aggregates / inclusive prefixes of the tiles before
// them, a warp at a time.  One read and one write per item whatever the length: 31 M carriers at GRCh38 size are one
// launch, where the previous single-block kernels walked the array 1024 elements at a time.
#ifndef MC_SCAN_CUH
#define MC_SCAN_CUH

#define MC_SCAN_TILE 2048   // 256 threads x 8 items
#define MC_SCAN_THREADS 256

struct ScanSum { static __device__ __forceinline__ int64_t op(int64_t a, int64_t b) { return a + b; } static __device__ __forceinline__ int64_t ident() { return 0; } };
struct ScanMax { static __device__ __forceinline__ int64_t op(int64_t a, int64_t b) { return a > b ? a : b; } static __device__ __forceinline__ int64_t ident() { return INT64_MIN; } };

// scratch of one scan: [0] ticket, then per tile a status word (0 nothing, 1 aggregate, 2 inclusive prefix), the aggregate and the inclusive prefix
static size_t device_scan_scratch_bytes(int64_t n) { const int64_t nt = (n + MC_SCAN_TILE - 1) / MC_SCAN_TILE + 1; return (size_t)(64 + nt * 4 + 64 + nt * 16); }
struct ScanScratch { int* ticket; int* status; int64_t* agg; int64_t* incl; };
static ScanScratch scan_scratch_of(void* p, int64_t n)
{
	const int64_t nt = (n + MC_SCAN_TILE - 1) / MC_SCAN_TILE + 1;
	ScanScratch s; s.ticket = (int*)p; s.status = (int*)((uint8_t*)p + 64);
	s.agg = (int64_t*)((uint8_t*)p + 64 + ((nt * 4 + 63) & ~63ll)); s.incl = s.agg + nt;
	return s;
}
static size_t scan_scratch_clear_bytes(int64_t n) { const int64_t nt = (n + MC_SCAN_TILE - 1) / MC_SCAN_TILE + 1; return (size_t)(64 + nt * 4); }

static __device__ __forceinline__ int scan_ld_status(const int* p) { return *(const volatile int*)p; }
static __device__ __forceinline__ int64_t scan_ld_value(const int64_t* p) { return *(const volatile int64_t*)p; }

// TIn: uint32_t or int64_t.  INCLUSIVE: out[i] = op(in[0..i]), else op(in[0..i-1]).  total (optional) receives op over everything.
// in == out is allowed (a tile reads all of its items before it writes any).
template <class TIn, class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(MC_SCAN_THREADS) mc_lookback_scan_kernel(const TIn* __restrict__ in, int64_t* out, int64_t n, ScanScratch sc, int64_t* total)
{
	__shared__ int tile_s;
	__shared__ int64_t wagg[MC_SCAN_THREADS / 32];
	__shared__ int64_t prefix_s;
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) tile_s = atomicAdd(sc.ticket, 1);
	__syncthreads();
	const int tile = tile_s;
	const int64_t base = (int64_t)tile * MC_SCAN_TILE + (int64_t)threadIdx.x * 8;
	int64_t v[8];
	int64_t s = Op::ident();
#pragma unroll
	for (int k = 0; k < 8; k++) { v[k] = base + k < n ? (int64_t)in[base + k] : Op::ident(); s = Op::op(s, v[k]); }
	// inclusive scan of the thread totals inside the warp, warp totals through shared memory
	int64_t incl = s;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(full, incl, o); if (lane >= o) incl = Op::op(t, incl); }
	if (lane == 31) wagg[warp] = incl;
	__syncthreads();
	int64_t wpre = Op::ident(), tile_agg = Op::ident();
#pragma unroll
	for (int w = 0; w < MC_SCAN_THREADS / 32; w++) { if (w < warp) wpre = Op::op(wpre, wagg[w]); tile_agg = Op::op(tile_agg, wagg[w]); }
	if (warp == 0)
	{
		int64_t p = Op::ident();
		if (tile == 0)
		{
			if (lane == 0) { sc.incl[0] = tile_agg; __threadfence(); *(volatile int*)&sc.status[0] = 2; }
		}
		else
		{
			if (lane == 0) { sc.agg[tile] = tile_agg; __threadfence(); *(volatile int*)&sc.status[tile] = 1; }
			// look back 32 tiles at a time: everything up to (and including) the nearest tile that already knows its inclusive prefix
			for (int64_t j = tile - 1;; j -= 32)
			{
				const int64_t idx = j - lane;
				int st = idx >= 0 ? scan_ld_status(sc.status + idx) : 2;       // "tile -1" is an inclusive prefix of nothing
				while (__any_sync(full, st == 0)) { if (st == 0) st = scan_ld_status(sc.status + idx); }
				__threadfence();
				int64_t val = Op::ident();
				if (idx >= 0) val = st == 2 ? scan_ld_value(sc.incl + idx) : scan_ld_value(sc.agg + idx);
				const unsigned done = __ballot_sync(full, st == 2);
				const int first = done ? __ffs((int)done) - 1 : 31;
				int64_t c = lane <= first ? val : Op::ident();
#pragma unroll
				for (int o = 16; o; o >>= 1) c = Op::op(c, __shfl_xor_sync(full, c, o));
				p = Op::op(c, p);
				if (done) break;
			}
			if (lane == 0) { sc.incl[tile] = Op::op(p, tile_agg); __threadfence(); *(volatile int*)&sc.status[tile] = 2; }
		}
		if (lane == 0)
		{
			prefix_s = p;
			if (total && (int64_t)(tile + 1) * MC_SCAN_TILE >= n) *total = Op::op(p, tile_agg);   // the last tile knows everything
		}
	}
	__syncthreads();
	int64_t run = Op::op(prefix_s, Op::op(wpre, __shfl_up_sync(full, incl, 1)));
	if (lane == 0) run = Op::op(prefix_s, wpre);
#pragma unroll
	for (int k = 0; k < 8; k++)
	{
		if (INCLUSIVE) { run = Op::op(run, v[k]); if (base + k < n) out[base + k] = run; }
		else { if (base + k < n) out[base + k] = run; run = Op::op(run, v[k]); }
	}
}

template <class TIn, class Op, bool INCLUSIVE>
static void device_lookback_scan(const TIn* in, int64_t* out, int64_t n, void* scratch, int64_t* total, cudaStream_t s)
{
	if (n <= 0) { if (total) cudaMemsetAsync(total, 0, 8, s); return; }
	const int64_t nt = (n + MC_SCAN_TILE - 1) / MC_SCAN_TILE;
	cudaMemsetAsync(scratch, 0, scan_scratch_clear_bytes(n), s);
	mc_lookback_scan_kernel<TIn, Op, INCLUSIVE><<<(unsigned)nt, MC_SCAN_THREADS, 0, s>>>(in, out, n, scan_scratch_of(scratch, n), total);
}

#endif
