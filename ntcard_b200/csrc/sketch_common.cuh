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
	// gap seeds (-g, stRead ntcard.cpp:160-171): gap > 0 only with nK == 1
	uint32_t gap, gap_a31, gap_a33, pad_;   // a = (k - gap) / 2 ones on either side of the gap; a % 31, a % 33
	KTab gtab;                               // rolling constants of the gap-long inner window
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

// ---- gap seeds -------------------------------------------------------------------------------------------------
// NTMSM64 (nthash.hpp:620-678, one seed, one hash) XORs the contribution of every '0' position i of the seed out of
// the plain hashes: fs = fh ^ srol^(k-1-i) seed[c_i], rs = rh ^ srol^i seed[comp c_i], h = min(fs, rs).  The seed of
// ntcard.cpp:407-413 is a ones, gap zeros, a ones (k = 2a + gap), so the zero positions are the inner window
// [a, a + gap) and their contributions are the ntHash of that window rotated by a:
//   XOR_j srol^(k-1-a-j) seed[c_(a+j)] = srol^a( XOR_j srol^(gap-1-j) seed[c_(a+j)] ) = srol^a( NTF64 of the inner window )
//   XOR_j srol^(a+j) seed[comp c_(a+j)]                                              = srol^a( NTR64 of the inner window )
// i.e. a second rolling hash of length gap instead of 2*gap table lookups per k-mer.
__device__ __forceinline__ uint64_t srol_by(uint64_t v, uint32_t a31, uint32_t a33)
{
	uint64_t hi = v >> 33, lo = v & 0x1FFFFFFFFull;
	hi = ((hi << a31) | (hi >> (31u - a31))) & 0x7FFFFFFFull;
	lo = ((lo << a33) | (lo >> (33u - a33))) & 0x1FFFFFFFFull;
	return (hi << 33) | lo;
}

__device__ __forceinline__ uint32_t process_piece_gap(const uint32_t* __restrict__ b, uint32_t len, uint32_t a0, uint32_t k, uint32_t gap,
    uint32_t a31, uint32_t a33, const KTab& T, const KTab& G, uint32_t* __restrict__ ctr_k, uint32_t rBits, uint32_t sBits)
{
	if (len < k)
		return 0;
	const uint32_t ns = len - k + 1;
	if (a0 >= ns)
		return 0;
	const uint32_t e = min(a0 + PIECE_STARTS, ns);
	const uint32_t a = (k - gap) / 2;
	uint64_t fh = 0, rh = 0, fg = 0, rg = 0;
	for (uint32_t i = 0; i < k; i++)
		fh = srol(fh) ^ seed_of(base_at(b, a0 + i));
	for (uint32_t i = k; i-- > 0;)
		rh = srol(rh) ^ seed_of(3u - base_at(b, a0 + i));
	for (uint32_t i = 0; i < gap; i++)
		fg = srol(fg) ^ seed_of(base_at(b, a0 + a + i));
	for (uint32_t i = gap; i-- > 0;)
		rg = srol(rg) ^ seed_of(3u - base_at(b, a0 + a + i));
	{
		const uint64_t fs = fh ^ srol_by(fg, a31, a33), rs = rh ^ srol_by(rg, a31, a33);
		sample_and_count(rs < fs ? rs : fs, ctr_k, rBits, sBits);
	}
	BaseStream so, si, go, gi;
	so.open(b, a0);
	go.open(b, a0 + a);
	if (a0 + 1 < e) {
		si.open(b, a0 + k);
		gi.open(b, a0 + a + gap);
	}
	for (uint32_t j = a0 + 1; j < e; j++) {
		const uint32_t idx = si.next() | (so.next() << 2);
		fh = srol(fh) ^ T.xf[idx];
		rh = sror(rh ^ T.xr[idx]);
		const uint32_t gidx = gi.next() | (go.next() << 2);
		fg = srol(fg) ^ G.xf[gidx];
		rg = sror(rg ^ G.xr[gidx]);
		const uint64_t fs = fh ^ srol_by(fg, a31, a33), rs = rh ^ srol_by(rg, a31, a33);
		sample_and_count(rs < fs ? rs : fs, ctr_k, rBits, sBits);
	}
	return e - a0;
}

#endif

} // namespace ntc
