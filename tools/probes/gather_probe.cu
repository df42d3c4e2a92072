// Micro-benchmark: random 32-byte gathers from a table that does not fit L2, with the load flavours the seed kernel could
// use.  Reports GB/s of useful bytes; run under ncu for the dram__bytes_read of each flavour.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu && ./gather_probe [MiB] [l2_granularity]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE> __device__ __forceinline__ uint32_t load32(const uint32_t* p)
{
	uint32_t a, b, c, d, e, f, g, h;
	if (MODE == 0)      // two 128-bit read-only loads
	{
		asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
		asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 4));
	}
	else if (MODE == 1) // one 256-bit read-only load
		asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
	else if (MODE == 2) // one 256-bit plain load
		asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
	else if (MODE == 3) // two 128-bit plain loads
	{
		asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
		asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 4));
	}
	else if (MODE == 4) // 256-bit, no L1 allocation
		asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
	else                // two 128-bit, no L1 allocation, read-only
	{
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 4));
	}
	return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

// every thread runs a dependent chain of gathers (the next block index comes from the loaded data), like an FM-index walk
template <int MODE> __global__ void gather(const uint32_t* tab, uint64_t n_blocks, int steps, uint32_t* out)
{
	uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
	uint32_t acc = 0;
	for (int s = 0; s < steps; s++)
	{
		const uint64_t b = (x >> 11) % n_blocks;
		const uint32_t v = load32<MODE>(tab + b * 8);
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
	printf("%-44s %8.3f ms  %8.1f GB/s useful (32 B per gather)  %s\n", name, ms, (double)threads * steps * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
	const size_t mib = argc > 1 ? atol(argv[1]) : 2048;
	if (argc > 2) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[2])); printf("set L2 fetch granularity %s: %s\n", argv[2], cudaGetErrorString(e)); }
	size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity %zu, table %zu MiB\n", gran, mib);
	uint32_t *tab, *out; const uint64_t nb = mib * 1024 * 1024 / 32;
	cudaMalloc(&tab, nb * 32); cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
	cudaMemset(tab, 0x5A, nb * 32);
	run<0>("2 x ld.global.nc.v4", tab, nb, out);
	run<1>("ld.global.nc.v8", tab, nb, out);
	run<2>("ld.global.v8", tab, nb, out);
	run<3>("2 x ld.global.v4", tab, nb, out);
	run<4>("ld.global.L1::no_allocate.v8", tab, nb, out);
	run<5>("2 x ld.global.nc.L1::no_allocate.v4", tab, nb, out);
	return 0;
}
