// sketch_common.cuh -- device-side parameter block and small helpers shared by the sketch kernels.
#pragma once
#include <stdint.h>

#include "../../include/ntcard_b200.h"
#include "nthash_device.cuh"

namespace ntc {

// Per-k rolling constants.  With in/out the 2-bit codes entering/leaving the window
// (nthash.hpp:242-257, NTF64/NTR64 sliding forms):
//   fh' = srol(fh) ^ seed[in] ^ srol^k(seed[out])            -> xf[in | out<<2]
//   rh' = sror(rh ^ srol^k(seed[comp in]) ^ seed[comp out])  -> xr[in | out<<2]
struct KTab {
	uint64_t xf[16];
	uint64_t xr[16];
};

struct DevParams {
	uint32_t nK, rBits, sBits, kmin;
	uint32_t k[NTC_MAX_K];
	KTab tab[NTC_MAX_K];
};

// k-mer starts handled by one thread of the general kernel ("piece").  Long records are cut into
// pieces of PIECE_STARTS starts; a piece needs PIECE_STARTS + k - 1 bases, so consecutive pieces
// overlap by k-1 bases (the hash is a pure function of the window, so this is exact).
constexpr uint32_t PIECE_STARTS = 512;

__device__ __forceinline__ uint32_t base_at(const uint32_t* __restrict__ b, uint32_t idx)
{
	return (__ldg(b + (idx >> 4)) >> ((idx & 15u) * 2u)) & 3u;
}

// ntComp (ntcard.cpp:132-145) on the device: counters are uint32 in HBM, [table][bucket] for one k;
// they are narrowed mod 2^16 at ntc_finish, which reproduces the reference's uint16_t wrap exactly.
__device__ __forceinline__ void sample_and_count(uint64_t h, uint32_t* __restrict__ ctr_k, uint32_t rBits, uint32_t sBits)
{
	const unsigned t = sample_table(h, sBits);
	if (t < 2u) {
		const uint64_t idx = ((uint64_t)t << rBits) + (h & (((uint64_t)1 << rBits) - 1));
		atomicAdd(ctr_k + idx, 1u); // result unused -> RED.E.ADD, fire and forget
	}
}

#if defined(__CUDACC__)
// ---- the general per-piece path (used by roll64_kernel and as the in-kernel fallback of the bit-sliced kernel) ----
struct BaseStream { // sequential 2-bit reader with lazy word refill (never reads past the last needed word)
	const uint32_t* __restrict__ p;
	uint32_t w;
	uint32_t left;
	__device__ __forceinline__ void open(const uint32_t* __restrict__ b, uint32_t idx)
	{
		p = b + (idx >> 4);
		w = __ldg(p) >> ((idx & 15u) * 2u);
		left = 16u - (idx & 15u);
	}
	__device__ __forceinline__ uint32_t next()
	{
		if (left == 0) {
			w = __ldg(++p);
			left = 16;
		}
		uint32_t c = w & 3u;
		w >>= 2;
		--left;
		return c;
	}
};

__device__ __forceinline__ uint32_t process_piece_k(const uint32_t* __restrict__ b, uint32_t len, uint32_t a, uint32_t k,
    const KTab& T, uint32_t* __restrict__ ctr_k, uint32_t rBits, uint32_t sBits)
{
	if (len < k)
		return 0;
	const uint32_t ns = len - k + 1;
	if (a >= ns)
		return 0;
	const uint32_t e = min(a + PIECE_STARTS, ns);
	// from-scratch hashes of the first window: NTF64/NTR64 base forms, nthash.hpp:220-239
	uint64_t fh = 0, rh = 0;
	{
		BaseStream s;
		s.open(b, a);
		for (uint32_t i = 0; i < k; i++)
			fh = srol(fh) ^ seed_of(s.next());
		for (uint32_t i = k; i-- > 0;)
			rh = srol(rh) ^ seed_of(3u - base_at(b, a + i));
	}
	sample_and_count(rh < fh ? rh : fh, ctr_k, rBits, sBits);
	BaseStream so, si;
	so.open(b, a);
	if (a + 1 < e)
		si.open(b, a + k);
	for (uint32_t j = a + 1; j < e; j++) {
		const uint32_t idx = si.next() | (so.next() << 2);
		fh = srol(fh) ^ T.xf[idx];
		rh = sror(rh ^ T.xr[idx]);
		sample_and_count(rh < fh ? rh : fh, ctr_k, rBits, sBits);
	}
	return e - a;
}

#endif

} // namespace ntc
