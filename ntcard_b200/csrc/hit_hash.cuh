// hit_hash.cuh -- the full 64-bit canonical ntHash of one sampled k-mer from the packed bases, and ntComp
// (ntcard.cpp:132-145) on it: shared by the fused sketch kernel (fused_kernel.cuh) and the stand-alone hit kernel
// (hit_kernels.cu).  Byte-indexed tables in shared memory: 8 x 128-bit lookups per 32 bases.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nthash_device.cuh"
#include "pipeline.h"

namespace ntc {
namespace pl {

constexpr size_t kTabBytes = 8 * 256 * 16;

// ------------------------------------------------------------------------------------------------
// full 64-bit canonical hash of one k-mer from the packed bases
// ------------------------------------------------------------------------------------------------
// srol applied n times, for n given as (a, b) = (n % 31, n % 33): rotate the upper ring by a, the lower by b.
__device__ __forceinline__ uint64_t srol_ab(uint64_t v, uint32_t a, uint32_t b)
{
	uint32_t hi = (uint32_t)(v >> 33);
	uint64_t lo = v & 0x1FFFFFFFFull;
	hi = ((hi << a) | (hi >> (31u - a))) & 0x7FFFFFFFu;
	lo = ((lo << b) | (lo >> (33u - b))) & 0x1FFFFFFFFull;
	return ((uint64_t)hi << 33) | lo;
}

struct HashCtx {
	const uint32_t* __restrict__ words;
	uint32_t stride, k, rBits, sBits;
	const uint4* tab; // shared: [8][256] {FB.lo, FB.hi, RB.lo, RB.hi}
	uint64_t rot_a, rot_b;
};

struct HitLoad { // the packed words of a candidate's first 32-base block, in flight
	uint32_t x0, x1, x2;
};

// rec_words: the base words of the record (word 0 = length skipped); p: k-mer start; last: index of the last base word
// that may be read (the record's slot in the uniform-stride batch)
template <bool kStaged> __device__ __forceinline__ uint32_t ld_word(const uint32_t* p)
{
	return kStaged ? *p : __ldg(p); // staged: the tile's packed reads sit in shared memory
}

template <bool kStaged>
__device__ __forceinline__ HitLoad hit_issue(const HashCtx& c, const uint32_t* __restrict__ rec_words, uint32_t p, uint32_t last)
{
	HitLoad h;
	const uint32_t wi = (p + (c.k & 31u)) >> 4;
	h.x0 = ld_word<kStaged>(rec_words + min(wi, last));
	h.x1 = ld_word<kStaged>(rec_words + min(wi + 1, last));
	h.x2 = ld_word<kStaged>(rec_words + min(wi + 2, last));
	return h;
}

// 32 bases (two packed words) through the byte tables: FB = XOR_i srol^(31-i) seed[c_i], RB = XOR_i srol^i seed[3-c_i]
__device__ __forceinline__ void block_tables(const uint4* __restrict__ tab, uint32_t w0, uint32_t w1, uint32_t& f0, uint32_t& f1,
    uint32_t& r0, uint32_t& r1)
{
	f0 = f1 = r0 = r1 = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const uint32_t w = j < 4 ? w0 : w1;
		const int sh = 8 * (j & 3) - 4; // byte j scaled by 16 (the entry size)
		const uint32_t off = (sh < 0 ? (w << 4) : (w >> sh)) & 0xFF0u;
		const uint4 e = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(tab) + j * 4096 + off);
		f0 ^= e.x;
		f1 ^= e.y;
		r0 ^= e.z;
		r1 ^= e.w;
	}
}

// Canonical hash of the k-mer and ntComp (ntcard.cpp:132-145).  k = t + 32*M: the t head bases one at a time
// (NTF64/NTR64 base forms, nthash.hpp:220-239), then M blocks of 32 bases:
//   fh = srol^32(fh) ^ FB_m        rh ^= srol^(t+32m) RB_m
// Returns the counter index inside the k's [2][2^rBits] sub-sketch, or kVoid when ntComp does not sample it.
template <bool kStaged>
__device__ __forceinline__ uint32_t hit_finish(const HashCtx& c, const HitLoad& h, const uint32_t* __restrict__ rec_words, uint32_t p, uint32_t last)
{
	const uint32_t t = c.k & 31u, M = c.k >> 5;
	uint32_t hh, hl; // canonical hash, high / low word
	if (t == 0 && M == 1) { // k = 32: pure 32-bit path
		const uint32_t sh = (p & 15u) * 2u;
		uint32_t f0, f1, r0, r1;
		block_tables(c.tab, __funnelshift_r(h.x0, h.x1, sh), __funnelshift_r(h.x1, h.x2, sh), f0, f1, r0, r1);
		const bool rlt = r1 < f1 || (r1 == f1 && r0 < f0);
		hh = rlt ? r1 : f1;
		hl = rlt ? r0 : f0;
	} else {
		uint64_t fh = 0, rh = 0;
		for (uint32_t i = 0; i < t; i++) {
			const uint32_t code = (ld_word<kStaged>(rec_words + ((p + i) >> 4)) >> (((p + i) & 15u) * 2u)) & 3u;
			fh = srol(fh) ^ seed_of(code);
			rh ^= srol_n(seed_of(3u - code), i);
		}
		uint32_t x0 = h.x0, x1 = h.x1, x2 = h.x2;
		for (uint32_t m = 0; m < M; m++) {
			const uint32_t o = p + t + 32u * m, sh = (o & 15u) * 2u;
			if (m) {
				const uint32_t wi = o >> 4;
				x0 = ld_word<kStaged>(rec_words + wi);
				x1 = ld_word<kStaged>(rec_words + wi + 1);
				x2 = ld_word<kStaged>(rec_words + min(wi + 2, last));
			}
			uint32_t f0, f1, r0, r1;
			block_tables(c.tab, __funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), f0, f1, r0, r1);
			const uint64_t FB = ((uint64_t)f1 << 32) | f0, RB = ((uint64_t)r1 << 32) | r0;
			fh = (m || t) ? (srol_ab(fh, 1, 32) ^ FB) : FB;
			const uint32_t ra = (uint32_t)(c.rot_a >> (8 * m)) & 0xFFu, rb = (uint32_t)(c.rot_b >> (8 * m)) & 0xFFu;
			rh ^= (ra | rb) ? srol_ab(RB, ra, rb) : RB;
		}
		const uint64_t hm = rh < fh ? rh : fh;
		hh = (uint32_t)(hm >> 32);
		hl = (uint32_t)hm;
	}
	// ntComp: both tests look at the top S+1 <= 32 bits; the bucket at the low rBits <= 30 bits
	const uint32_t S = c.sBits;
	const bool t0 = (hh >> (31 - S)) == 1u;
	const bool t1 = (hh >> (32 - S)) == ((1u << (S - 1)) - 1u);
	if (!(t0 || t1))
		return kVoid;
	return ((t1 ? 1u : 0u) << c.rBits) | (hl & ((1u << c.rBits) - 1u));
}

} // namespace pl
} // namespace ntc
