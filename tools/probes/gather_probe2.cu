// Micro-benchmark, second round: can a random 32-byte gather from HBM cost less than a whole 128-byte line?  gather_probe.cu
// showed ~123 B of DRAM traffic per gather for ld.global[.nc][.L1::no_allocate] v4 / v8.  Here: L2-only loads (.cg), the
// prefetch-size hints, an evict-first policy, asynchronous copies into shared memory that bypass L1 (cp.async.cg) and the
// bulk-copy engine (cp.async.bulk, 32 bytes per copy) - each as a dependent chain per thread, like an FM-index walk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe2 gather_probe2.cu
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum ./gather_probe2 [MiB]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE> __device__ __forceinline__ uint32_t load32(const uint32_t* p, uint32_t* sm, uint64_t* bar, uint64_t pol, uint32_t& phase)
{
	uint32_t a = 0, b = 0, c = 0, d = 0, e = 0, f = 0, g = 0, h = 0;
	if (MODE == 0)
		asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
	else if (MODE == 1)
	{
		asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
		asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 4));
	}
	else if (MODE == 2)
		asm volatile("ld.global.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
	else if (MODE == 3)
		asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p), "l"(pol));
	else if (MODE == 4)   // 2 x 16-byte asynchronous copies into shared memory, L2 only
	{
		const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
		asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(p));
		asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s + 16), "l"(p + 4));
		asm volatile("cp.async.wait_all;" ::: "memory");
		a = sm[0]; b = sm[1]; c = sm[2]; d = sm[3]; e = sm[4]; f = sm[5]; g = sm[6]; h = sm[7];
	}
	else if (MODE == 5)   // bulk-copy engine: 32 bytes per lane onto a per-warp mbarrier
	{
		const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm), mb = (uint32_t)__cvta_generic_to_shared(bar);
		if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" :: "r"(mb), "r"(32 * 32));
		__syncwarp();
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" :: "r"(s), "l"(p), "r"(mb) : "memory");
		uint32_t ok = 0;
		while (!ok) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(ok) : "r"(mb), "r"(phase) : "memory");
		phase ^= 1;
		a = sm[0]; b = sm[1]; c = sm[2]; d = sm[3]; e = sm[4]; f = sm[5]; g = sm[6]; h = sm[7];
		__syncwarp();
	}
	else if (MODE == 6)   // only 8 of the 32 bytes
	{
		asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
	}
	else                  // volatile-style uncached
		asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
	return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int MODE> __global__ void __launch_bounds__(256) gather(const uint32_t* tab, uint64_t n_blocks, int steps, uint32_t* out)
{
	__shared__ __align__(128) uint32_t sm[256 * 8];
	__shared__ __align__(8) uint64_t bars[8];
	uint64_t pol = 0;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	if (MODE == 5 && (threadIdx.x & 31) == 0) asm volatile("mbarrier.init.shared.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bars[threadIdx.x >> 5])));
	__syncthreads();
	uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
	uint32_t acc = 0, phase = 0;
	for (int s = 0; s < steps; s++)
	{
		const uint64_t b = (x >> 11) % n_blocks;
		const uint32_t v = load32<MODE>(tab + b * 8, sm + threadIdx.x * 8, &bars[threadIdx.x >> 5], pol, phase);
		acc ^= v;
		x = x * 6364136223846793005ull + 1442695040888963407ull + v;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> static void run(const char* name, const uint32_t* tab, uint64_t nb, uint32_t* out)
{
	const int steps = 64, threads = 148 * 8 * 256 * 4;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	gather<MODE><<<threads / 256, 256>>>(tab, nb, steps, out);
	cudaEventRecord(e0);
	gather<MODE><<<threads / 256, 256>>>(tab, nb, steps, out);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	printf("%-52s %8.3f ms  %8.1f GB/s useful (32 B per gather)  %s\n", name, ms, (double)threads * steps * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
	const size_t mib = argc > 1 ? atol(argv[1]) : 2048;
	uint32_t *tab, *out; const uint64_t nb = mib * 1024 * 1024 / 32;
	cudaMalloc(&tab, nb * 32); cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
	cudaMemset(tab, 0x5A, nb * 32);
	printf("table %zu MiB, %d gathers\n", mib, 148 * 8 * 256 * 4 * 64);
	run<0>("ld.global.nc.v8 (baseline)", tab, nb, out);
	run<1>("2 x ld.global.cg.v4", tab, nb, out);
	run<2>("ld.global.L2::64B.v8", tab, nb, out);
	run<3>("ld.global.nc.L1::no_allocate.L2::cache_hint(evict_first).v8", tab, nb, out);
	run<4>("2 x cp.async.cg 16 B -> shared", tab, nb, out);
	run<5>("cp.async.bulk 32 B -> shared (mbarrier per warp)", tab, nb, out);
	run<6>("ld.global.nc.v2 (8 of the 32 bytes)", tab, nb, out);
	run<7>("ld.global.cv.v4 (16 of the 32 bytes)", tab, nb, out);
	return 0;
}
