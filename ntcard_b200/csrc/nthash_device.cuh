// nthash_device.cuh -- ntHash primitives shared by host setup code and sm_100a kernels.
//
// What is computed is fixed by the reference (bcgsc/ntCard v1.2.2, vendor/ntHash/nthash.hpp);
// how it is computed is ours.  Citations are reference file:line.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NTC_HD __host__ __device__ __forceinline__
#else
#define NTC_HD inline
#endif

namespace ntc {

// 2-bit base codes of the packed record stream: A=0 C=1 G=2 T/U=3; complement = 3 - code.
// Seeds: nthash.hpp:25-28 (seedA, seedC, seedG, seedT).
NTC_HD uint64_t seed_of(unsigned code)
{
	return code == 0 ? 0x3c8bfbb395c60474ULL
	     : code == 1 ? 0x3193c18562a02b4cULL
	     : code == 2 ? 0x20323ed082572324ULL
	                 : 0x295549f54be24456ULL;
}

// srol: the 64-bit word is two independent rings, bits 63..33 (31 bits) and bits 32..0 (33 bits),
// each rotated left by one: swapbits033(rol1(v)), nthash.hpp:186-188,208-211.
NTC_HD uint64_t srol(uint64_t v)
{
	uint64_t r = (v << 1) | (v >> 63);
	uint64_t x = (r ^ (r >> 33)) & 1;
	return r ^ (x | (x << 33));
}

// inverse: swapbits3263(ror1(v)), nthash.hpp:191-193,214-217.
NTC_HD uint64_t sror(uint64_t v)
{
	uint64_t r = (v >> 1) | (v << 63);
	uint64_t x = ((r >> 32) ^ (r >> 63)) & 1;
	return r ^ ((x << 32) | (x << 63));
}

// srol applied n times == msTab31l[c][n%31] | msTab33r[c][n%33] for a seed (nthash.hpp:66-183,196-205).
NTC_HD uint64_t srol_n(uint64_t v, unsigned n)
{
	uint64_t hi = v >> 33, lo = v & 0x1FFFFFFFFULL;
	unsigned a = n % 31, b = n % 33;
	if (a) hi = ((hi << a) | (hi >> (31 - a))) & 0x7FFFFFFFULL;
	if (b) lo = ((lo << b) | (lo >> (33 - b))) & 0x1FFFFFFFFULL;
	return (hi << 33) | lo;
}

// ntComp's sampling rule, ntcard.cpp:135-139: table 0 when the top sBits+1 bits are 0..01,
// table 1 when the top sBits bits are 01..1 (the later test wins), else 2 = not sampled.
NTC_HD unsigned sample_table(uint64_t h, unsigned sBits)
{
	const uint64_t sMask = ((uint64_t)1 << (sBits - 1)) - 1; // ntcard.cpp:438
	unsigned t = 2;
	if ((h >> (63 - sBits)) == 1) t = 0;
	if ((h >> (64 - sBits)) == sMask) t = 1;
	return t;
}

} // namespace ntc
