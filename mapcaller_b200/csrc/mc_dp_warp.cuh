// Warp-per-task gapped fill (CUDA only) for everything that is not "small" (dp_is_small, mc_stages_align.h: those go one
// thread per fill through dp_core in shared memory, mc_dp_small_kernel below).
//
// One warp fills one (read piece x genome piece) matrix as an anti-diagonal wavefront: lane l owns one column of a
// 32-column strip, at step t it computes row t-l, the left / diagonal neighbours arrive by __shfl_up_sync from lane
// l-1, the values of the last column of a strip are parked in the task's workspace for lane 0 of the next strip.
// The three-way maxima use the Blackwell DPX instructions (__vimax3_s32 / __viaddmax_s32).  One traceback byte per
// cell is stored anti-diagonal-major (32 contiguous bytes per step), lane 0 then walks it back exactly as the
// reference does and rewrites the two strings in place.
//   nw   : reference src/nw_alignment.cpp:18-83 in doubled integers; rows = read piece, columns = genome piece
//   ksw2 : reference src/ksw2_alignment.cpp:70-272;                    rows = genome piece (target), columns = read piece (query)
// dp_core of mc_stages_align.h computes the same thing serially (small fills, and the developer harness); both are checked against the oracle by tests/test_parity_gpu.py::test_gapped_fill_kernel_matches_oracle.
#ifndef MC_DP_WARP_CUH
#define MC_DP_WARP_CUH

#define MC_NW_NEG (-131072)

template <bool KSW>
static __device__ __forceinline__ void dpw_task(const PipeArgs& a, const DpTask tk, const int lane, uint32_t& cells, uint8_t* fast, int fast_bytes)
{
	const unsigned full = 0xffffffffu;
	mc_frag_out& x = a.frags[tk.frag];
	uint8_t* s1 = a.aln + x.aln_off; uint8_t* s2 = s1 + x.aln_cap;     // read piece (m), genome piece (n)
	const int m = tk.m, n = tk.n;
	const uint8_t* rowstr = KSW ? s2 : s1; const uint8_t* colstr = KSW ? s1 : s2;
	const int R = KSW ? n : m, C = KSW ? m : n;
	const int nstrips = (C + 31) >> 5;
	const int64_t strip_bytes = (int64_t)(R + 32) * 32;
	uint8_t* gws = a.dpws + tk.ws_off;
	int* bs = (int*)(gws + dp_tb_bytes(m, n)); int* bx = bs + ((m > n ? m : n) + 2);
	// the traceback bytes of a small fill (almost all of them) stay in shared memory: the walk back is a chain of dependent loads
	uint8_t* tb = (int64_t)nstrips * strip_bytes <= fast_bytes ? fast : gws;
	for (int strip = 0; strip < nstrips; strip++)
	{
		const int j = strip * 32 + lane;                 // column, 0-based
		const bool colok = j < C;
		const int cj = colok ? mc_nt4(colstr[j]) : 5;
		const int top = KSW ? -(2 + (j + 1)) : -2 - (j + 1);   // border above row 0 (same value in both scorings)
		int curS = top, curX = KSW ? MC_NEG_INF : top;   // this column at the row just computed: s / H and r / F
		int upS = top, upY = KSW ? MC_NEG_INF : MC_NW_NEG;     // same column one row up: s / H and t / E
		int diag = __shfl_up_sync(full, curS, 1);
		if (lane == 0) diag = strip == 0 ? 0 : bs[0];
		int rowc = 5, rowchunk = 5;
		uint8_t* tbs = tb + strip * strip_bytes;
		const int steps = R + 31;
		for (int t = 0; t < steps; t++)
		{
			const int i = t - lane;                      // row, 0-based
			if ((t & 31) == 0) rowchunk = t + lane < R ? mc_nt4(rowstr[t + lane]) : 5;   // 32 row characters per load, handed out by shuffle
			const int rc_new = __shfl_sync(full, rowchunk, t & 31);
			const int rc_recv = __shfl_up_sync(full, rowc, 1);
			rowc = lane == 0 ? rc_new : rc_recv;
			int leftS = __shfl_up_sync(full, curS, 1), leftX = __shfl_up_sync(full, curX, 1);
			if (lane == 0)
			{
				if (strip == 0) { leftS = KSW ? -(2 + (i + 1)) : -2 - (i + 1); leftX = KSW ? MC_NEG_INF : MC_NW_NEG; }
				else if (i < R) { leftS = bs[i + 1]; leftX = bx[i + 1]; }
			}
			const bool valid = colok && i >= 0 && i < R;
			int nS, nX, nY; uint8_t d;
			if (!KSW)
			{
				nX = __viaddmax_s32(leftX, -1, leftS - 3);             // r[i][j]
				nY = __viaddmax_s32(upY, -1, upS - 3);                 // t[i][j]
				const int dd = diag + (rowc == cj ? 2 : -2);
				nS = __vimax3_s32(dd, nX, nY);
				d = (uint8_t)((nS == nX ? 1 : 0) | (nS == nY ? 2 : 0));
			}
			else
			{
				const int sc = (rowc == 4 || cj == 4) ? 0 : (rowc == cj ? 1 : -1);
				nY = i > 0 ? __viaddmax_s32(upS, -2, upY) - 1 : MC_NEG_INF;        // E(i,j)
				nX = j > 0 ? __viaddmax_s32(leftS, -2, leftX) - 1 : MC_NEG_INF;     // F(i,j)
				int z = diag + sc; d = 0;
				if (nY > z) { d = 1; z = nY; }
				if (nX > z) { d = 2; z = nX; }
				if (nY > z - 2) d |= 0x08;
				if (nX > z - 2) d |= 0x10;
				nS = z;
			}
			diag = leftS;                                // s[i][j-1] is the diagonal of the next row
			if (valid)
			{
				curS = nS; curX = nX; upS = nS; upY = nY;
				tbs[t * 32 + lane] = d;
				if (lane == 31 && strip + 1 < nstrips) { bs[i + 1] = nS; bx[i + 1] = nX; }
			}
			__syncwarp();
		}
		if (lane == 31 && strip + 1 < nstrips) bs[0] = top;
		__syncwarp();
	}
	__syncwarp();
	if (lane == 0)
	{
		// one pass from the end of the matrix, writing the gapped strings right-aligned into their capacity (the write cursor
		// never passes an unread source byte: k >= i + j - 1 throughout); aln_off then moves to where they start
		const int cap = x.aln_cap;
		int k = cap;
		if (!KSW)
		{
			int i = m, j = n;
			while (i > 0 || j > 0)
			{
				const uint8_t d = i == 0 ? 1 : j == 0 ? 2 : tb[((j - 1) >> 5) * strip_bytes + (int64_t)(i - 1 + ((j - 1) & 31)) * 32 + ((j - 1) & 31)];
				k--;
				if (d & 1) { s2[k] = s2[j - 1]; s1[k] = '-'; j--; }
				else if (d & 2) { s1[k] = s1[i - 1]; s2[k] = '-'; i--; }
				else { s1[k] = s1[i - 1]; s2[k] = s2[j - 1]; i--; j--; }
			}
		}
		else
		{
			int i = n - 1, j = m - 1, state = 0;
			while (i >= 0 && j >= 0)
			{
				const uint32_t tmp = tb[(j >> 5) * strip_bytes + (int64_t)(i + (j & 31)) * 32 + (j & 31)];
				if (state == 0) state = tmp & 7;
				else if (!((tmp >> (state + 2)) & 1)) state = 0;
				if (state == 0) state = tmp & 7;
				k--;
				if (state == 0) { s1[k] = s1[j]; s2[k] = s2[i]; i--; j--; }
				else if (state == 1 || state == 3) { s2[k] = s2[i]; s1[k] = '-'; i--; }
				else { s1[k] = s1[j]; s2[k] = '-'; j--; }
			}
			while (i >= 0) { k--; s2[k] = s2[i]; s1[k] = '-'; i--; }
			while (j >= 0) { k--; s1[k] = s1[j]; s2[k] = '-'; j--; }
		}
		x.aln_off += k; x.aln_len = cap - k;
		cells += (uint32_t)(m * n);
	}
}

// persistent warps: warp w takes tasks task_begin + w, + n_warps, ...
#define MC_DP_SMEM (6 * 1024)     // traceback bytes per warp kept on chip: 6 KB covers fills up to ~160 rows x 32 columns
__global__ void __launch_bounds__(MC_BLOCK) mc_dp_kernel(const PipeArgs a)
{
	__shared__ __align__(16) uint8_t dp_smem[(MC_BLOCK / 32) * MC_DP_SMEM];
	uint8_t* fast = dp_smem + (threadIdx.x >> 5) * MC_DP_SMEM;
	const int lane = threadIdx.x & 31;
	const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	if (a.st->overflow) return;        // an arena ran out earlier in this attempt (a task slot may be unwritten): the attempt is repeated
	int64_t end = (int64_t)*a.task_bump; if (end > a.task_cap) end = a.task_cap;
	uint32_t cells = 0, tasks = 0;     // lane 0 of the warp counts, one atomic per warp at the end
	for (int64_t t = a.task_begin + warp; t < end; t += n_warps)
	{
		const DpTask tk = a.tasks[t];
		if (dp_is_small(tk.m, tk.n)) continue;            // mc_dp_small_kernel's
		if (a.pr.alg_ksw2) dpw_task<true>(a, tk, lane, cells, fast, MC_DP_SMEM); else dpw_task<false>(a, tk, lane, cells, fast, MC_DP_SMEM);
		tasks++;
		__syncwarp();
	}
	if (lane == 0 && tasks) { atomicAdd(&a.st->dp_cells, (mc_u64)cells); atomicAdd(&a.st->dp_tasks, (mc_u64)tasks); }
}
// small fills: persistent threads, thread i takes tasks task_begin + i, + n_threads, ... and skips the large ones
#define MC_DP_SMALL_THREADS 128
__global__ void __launch_bounds__(MC_DP_SMALL_THREADS) mc_dp_small_kernel(const PipeArgs a)
{
	extern __shared__ __align__(16) uint8_t dp_small_smem[];
	uint8_t* ws = dp_small_smem + (size_t)threadIdx.x * MC_DP_SMALL_STRIDE;
	const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, n_threads = (int64_t)gridDim.x * blockDim.x;
	if (a.st->overflow) return;        // see mc_dp_kernel
	int64_t end = (int64_t)*a.task_bump; if (end > a.task_cap) end = a.task_cap;
	uint32_t cells = 0, tasks = 0;
	for (int64_t t = a.task_begin + tid; t < end; t += n_threads) dp_small_body(t, a, ws, &cells, &tasks);
	__syncwarp();
	mc_stat_add(&a.st->dp_cells, cells); mc_stat_add(&a.st->dp_tasks, tasks);
}
static void launch_dp(const PipeArgs& a, int64_t max_tasks, mc_stream_t s)
{
	if (max_tasks <= 0) return;
	static bool configured[64];          // per device, see launch_rescue
	const int smem = MC_DP_SMALL_THREADS * MC_DP_SMALL_STRIDE;
	int dev = 0; cudaGetDevice(&dev); dev &= 63;
	if (!configured[dev]) { cudaFuncSetAttribute(mc_dp_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); configured[dev] = true; }
	int64_t sb = (max_tasks + MC_DP_SMALL_THREADS - 1) / MC_DP_SMALL_THREADS; if (sb > 148 * 2) sb = 148 * 2;
	mc_dp_small_kernel<<<(unsigned)sb, MC_DP_SMALL_THREADS, smem, s>>>(a); g_launches++;
	int64_t blocks = (max_tasks * 32 + MC_BLOCK - 1) / MC_BLOCK; if (blocks > 148 * 8) blocks = 148 * 8;
	mc_dp_kernel<<<(unsigned)blocks, MC_BLOCK, 0, s>>>(a); g_launches++;
}

#endif
