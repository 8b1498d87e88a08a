// bitslice_dispatch.cu -- picks the compiled (k mod 31) variant of the scan kernel; builds the byte tables of the
// full hash.  BS_KM_LIST is the list of compiled variants, e.g. -DBS_KM_LIST="X(0) X(1) X(2)".
#include "nthash_device.cuh"
#include "pipeline.h"

#ifndef BS_KM_LIST
#define BS_KM_LIST
#endif

namespace ntc {
namespace pl {

#define X(n) cudaError_t launch_scan_km_##n(unsigned sBits, const ScanArgs& a); cudaError_t launch_fused_km_##n(unsigned sBits, const FusedArgs& a); cudaError_t launch_hllscan_km_##n(unsigned T, const ScanArgs& a);
BS_KM_LIST
#undef X

bool have_scan_kernel(unsigned k, unsigned sBits)
{
	if (sBits != 7 && sBits != 11)
		return false;
	switch (k % 31) {
#define X(n) case n: return true;
		BS_KM_LIST
#undef X
	default: return false;
	}
}

cudaError_t launch_scan(unsigned k, unsigned sBits, const ScanArgs& a)
{
	switch (k % 31) {
#define X(n) case n: return launch_scan_km_##n(sBits, a);
		BS_KM_LIST
#undef X
	default: return cudaErrorInvalidValue;
	}
}

cudaError_t launch_hllscan(unsigned k, unsigned T, const ScanArgs& a)
{
	switch (k % 31) {
#define X(n) case n: return launch_hllscan_km_##n(T, a);
		BS_KM_LIST
#undef X
	default: return cudaErrorInvalidValue;
	}
}

cudaError_t launch_fused(unsigned k, unsigned sBits, const FusedArgs& a)
{
	switch (k % 31) {
#define X(n) case n: return launch_fused_km_##n(sBits, a);
		BS_KM_LIST
#undef X
	default: return cudaErrorInvalidValue;
	}
}

// shared memory of the fused kernel: byte tables + per warp (plane ring + 3 mirror slots, candidate queue, log state); fused_kernel.cuh
size_t fused_smem_bytes(uint32_t ring, uint32_t nwarps, uint32_t qlane, uint32_t nbins)
{
	const size_t queue = (((size_t)qlane * 64) + 15) & ~(size_t)15;
	return 8 * 256 * 16 + (size_t)nwarps * ((size_t)(ring + 3u) * 256u + queue + 7 * (size_t)nbins * 4u);
}

// Per byte position j (4 bases) of a 32-base block: FB = XOR_u srol^(31-i) seed[c_i], RB = XOR_u srol^i seed[3-c_i], i = 4j+u
void build_tables(uint32_t* tab)
{
	for (unsigned j = 0; j < 8; j++)
		for (unsigned b = 0; b < 256; b++) {
			uint64_t fb = 0, rb = 0;
			for (unsigned u = 0; u < 4; u++) {
				const unsigned code = (b >> (2 * u)) & 3, i = 4 * j + u;
				fb ^= srol_n(seed_of(code), 31 - i);
				rb ^= srol_n(seed_of(3 - code), i);
			}
			uint32_t* e = tab + (j * 256 + b) * 4;
			e[0] = (uint32_t)fb;
			e[1] = (uint32_t)(fb >> 32);
			e[2] = (uint32_t)rb;
			e[3] = (uint32_t)(rb >> 32);
		}
}

} // namespace pl
} // namespace ntc
