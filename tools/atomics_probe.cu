// tools/atomics_probe.cu -- microbenchmarks that decide how the sketch update is implemented.
// Measures on the B200: random RED.ADD throughput into HBM-sized vs L2-sized tables, a hit-log
// (append) write rate, LOP3 / IMAD issue rates, and 2^27-counter table streaming.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/atomics_probe tools/atomics_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ULL;
	uint64_t z = x;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

// each thread issues `per` random increments into table[0..mask]
__global__ void red_u32(uint32_t* table, uint64_t mask, int per, uint64_t salt)
{
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t h = mix64(t ^ salt);
	for (int i = 0; i < per; i++) {
		h = h * 6364136223846793005ULL + 1442695040888963407ULL;
		atomicAdd(table + ((h >> 20) & mask), 1u);
	}
}

// same but only one lane in `sparsity` does an atomic per iteration (models 1/64 sampling inside a busy warp)
__global__ void red_u32_sparse(uint32_t* table, uint64_t mask, int per, uint64_t salt, uint32_t keep_mask)
{
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t h = mix64(t ^ salt);
	for (int i = 0; i < per; i++) {
		h = h * 6364136223846793005ULL + 1442695040888963407ULL;
		if (((h >> 50) & keep_mask) == 0)
			atomicAdd(table + ((h >> 20) & mask), 1u);
	}
}

// 16-bit counters packed two per word
__global__ void red_u16pair(uint32_t* table, uint64_t mask, int per, uint64_t salt)
{
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t h = mix64(t ^ salt);
	for (int i = 0; i < per; i++) {
		h = h * 6364136223846793005ULL + 1442695040888963407ULL;
		uint64_t idx = (h >> 20) & mask;
		atomicAdd(table + (idx >> 1), 1u << (16 * (idx & 1)));
	}
}

// append log: every thread writes `per` words to a per-warp-coalesced stream
__global__ void log_write(uint32_t* log, int per)
{
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t n = (uint64_t)gridDim.x * blockDim.x;
	for (int i = 0; i < per; i++)
		log[(uint64_t)i * n + t] = (uint32_t)(t + i);
}

// binned scatter: each warp writes each value to one of `nbins` streams (position via per-bin global cursor, block-aggregated)
__global__ void binned_scatter(uint32_t* bins, uint32_t* cursors, uint32_t bin_cap, int nbins_log2, int per, uint64_t salt)
{
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t h = mix64(t ^ salt);
	for (int i = 0; i < per; i++) {
		h = h * 6364136223846793005ULL + 1442695040888963407ULL;
		uint32_t v = (uint32_t)(h >> 20) & 0x0FFFFFFF;
		uint32_t b = v >> (28 - nbins_log2);
		uint32_t pos = atomicAdd(cursors + b, 1u);
		if (pos < bin_cap)
			bins[(uint64_t)b * bin_cap + pos] = v;
	}
}

// apply: read a bin's values and increment within an L2-resident slice
__global__ void apply_bin(const uint32_t* vals, uint64_t n, uint32_t* table)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		atomicAdd(table + vals[i], 1u);
}

template <int ILP>
__global__ void lop3_rate(uint32_t* out, int iters)
{
	uint32_t a[ILP];
	uint32_t x = threadIdx.x * 2654435761u, y = blockIdx.x * 40503u + 1;
#pragma unroll
	for (int j = 0; j < ILP; j++) a[j] = x + j;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int j = 0; j < ILP; j++)
			asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(x), "r"(y));
		x += 1;
	}
	uint32_t s = 0;
#pragma unroll
	for (int j = 0; j < ILP; j++) s ^= a[j];
	if (s == 0x12345678) out[0] = s;
}

template <int ILP>
__global__ void mixed_rate(uint32_t* out, int iters)
{
	uint32_t a[ILP], b[ILP];
	uint32_t x = threadIdx.x * 2654435761u, y = blockIdx.x * 40503u + 1;
#pragma unroll
	for (int j = 0; j < ILP; j++) { a[j] = x + j; b[j] = y + j; }
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int j = 0; j < ILP; j++) {
			asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(x), "r"(y));
			asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[j]) : "r"(x), "r"(y));
		}
		x += 1;
	}
	uint32_t s = 0;
#pragma unroll
	for (int j = 0; j < ILP; j++) s ^= a[j] ^ b[j];
	if (s == 0x12345678) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps = 3)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	f(); // warm
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int r = 0; r < reps; r++) {
		cudaEventRecord(e0);
		f();
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	return best;
}

int main()
{
	cudaDeviceProp p;
	CK(cudaGetDeviceProperties(&p, 0));
	printf("device %s sm_%d%d SMs %d L2 %d MB clock %d MHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.l2CacheSize >> 20, p.clockRate / 1000);
	const int nsm = p.multiProcessorCount;
	uint32_t* table;
	const uint64_t big = 1ull << 28; // 2^28 u32 = 1 GiB
	CK(cudaMalloc(&table, big * 4));
	CK(cudaMemset(table, 0, big * 4));
	const int block = 256, grid = nsm * 16, per = 256;
	const double nat = (double)grid * block * per;
	for (int lg = 28; lg >= 16; lg -= 2) {
		uint64_t mask = (1ull << lg) - 1;
		float ms = time_ms([&] { red_u32<<<grid, block>>>(table, mask, per, lg); });
		printf("RED.u32 random into 2^%d counters (%6.1f MiB): %8.3f ms  %7.2f G atomics/s\n", lg, (double)(4ull << lg) / 1048576, ms, nat / ms / 1e6);
	}
	for (int lg : { 27, 25, 23 }) {
		uint64_t mask = (1ull << (lg + 1)) - 1; // lg+1 bits of u16 index -> 2^lg words
		float ms = time_ms([&] { red_u16pair<<<grid, block>>>(table, mask, per, lg); });
		printf("RED u16-pair random into 2^%d u16 counters (%6.1f MiB): %8.3f ms  %7.2f G atomics/s\n", lg + 1, (double)(2ull << (lg + 1)) / 1048576, ms, nat / ms / 1e6);
	}
	for (uint32_t keep : { 63u, 1023u }) {
		float ms = time_ms([&] { red_u32_sparse<<<grid, block>>>(table, big - 1, per * 16, 99, keep); });
		double n = nat * 16 / (keep + 1);
		printf("RED.u32 sparse 1/%u lanes into 1 GiB: %8.3f ms  %7.2f G atomics/s (%.1f G candidates/s)\n", keep + 1, ms, n / ms / 1e6, nat * 16 / ms / 1e6);
	}
	{
		uint32_t* log;
		CK(cudaMalloc(&log, (size_t)grid * block * 64 * 4));
		float ms = time_ms([&] { log_write<<<grid, block>>>(log, 64); });
		printf("hit-log coalesced append: %8.3f ms  %7.1f GB/s\n", ms, (double)grid * block * 64 * 4 / ms / 1e6);
		cudaFree(log);
	}
	for (int nb : { 3, 4, 5, 6 }) {
		uint32_t *bins, *cur;
		const uint32_t cap = (uint32_t)(nat / (1 << nb) * 1.1);
		CK(cudaMalloc(&bins, (size_t)cap * (1 << nb) * 4));
		CK(cudaMalloc(&cur, 4 << nb));
		float ms = time_ms([&] { cudaMemsetAsync(cur, 0, 4 << nb); binned_scatter<<<grid, block>>>(bins, cur, cap, nb, per, 7); });
		printf("binned scatter (naive per-element cursor atomics) into %d bins: %8.3f ms  %7.2f G elems/s\n", 1 << nb, ms, nat / ms / 1e6);
		// apply one bin into its slice
		uint64_t slice = 1ull << (28 - nb);
		uint32_t h_n;
		cudaMemcpy(&h_n, cur, 4, cudaMemcpyDeviceToHost);
		if (h_n > cap) h_n = cap;
		float ms2 = time_ms([&] { apply_bin<<<nsm * 8, 256>>>(bins, h_n, table); });
		printf("   apply bin 0 (%u values) into %5.1f MiB slice: %8.3f ms  %7.2f G atomics/s\n", h_n, (double)slice * 4 / 1048576, ms2, h_n / ms2 / 1e6);
		cudaFree(bins);
		cudaFree(cur);
	}
	{
		float ms = time_ms([&] { cudaMemsetAsync(table, 0, big * 4); });
		printf("memset 1 GiB: %8.3f ms  %7.1f GB/s\n", ms, (double)big * 4 / ms / 1e6);
	}
	{
		uint32_t* out;
		CK(cudaMalloc(&out, 4));
		const int iters = 4096;
		for (int warps : { 4, 8, 16, 32 }) {
			float ms = time_ms([&] { lop3_rate<8><<<nsm, warps * 32>>>(out, iters); });
			double ops = (double)nsm * warps * 32 * iters * 8;
			printf("LOP3 rate, %2d warps/SM ILP8: %7.2f T lane-ops/s  (%.1f lanes/clk/SM at %d MHz nominal)\n", warps, ops / ms / 1e9, ops / ms / 1e3 / nsm / (p.clockRate / 1000.0) , p.clockRate / 1000);
		}
		for (int warps : { 8, 16, 32 }) {
			float ms = time_ms([&] { mixed_rate<8><<<nsm, warps * 32>>>(out, iters); });
			double ops = (double)nsm * warps * 32 * iters * 16;
			printf("LOP3+IMAD rate, %2d warps/SM ILP8: %7.2f T lane-ops/s  (%.1f lanes/clk/SM nominal)\n", warps, ops / ms / 1e9, ops / ms / 1e3 / nsm / (p.clockRate / 1000.0));
		}
		cudaFree(out);
	}
	cudaFree(table);
	return 0;
}
