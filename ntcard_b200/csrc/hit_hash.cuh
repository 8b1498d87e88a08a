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

// ---- block-wise full hash with 128-bit gathers (fused kernel) ---------------------------------------------------------------
// k = tprime + 32 (nblk - 1), 1 <= tprime <= 32: the first tprime bases form the HEAD block, put through the byte tables with the
// other 32 - tprime codes masked to 0 ('A'); what those A's contribute is a constant per k (head_c / head_d), and the head's hash
// sits 32 - tprime rotations too far left:
//   fh = sror^(32-tprime)( FB(head) ^ head_c ),  rh = RB(head) ^ head_d;   then per full block m = 1 .. nblk-1:
//   fh = srol^32(fh) ^ FB_m,                      rh ^= srol^(tprime + 32 (m-1)) RB_m
// (NTF64 / NTR64 base forms, nthash.hpp:220-239, regrouped; checked against them for k = 1 .. 287 on the host.)
struct HashK {
	uint32_t k, tprime, nblk;
	uint32_t head_ra, head_rb;   // srol amounts (mod 31, mod 33) equal to sror^(32 - tprime)
	uint64_t head_c, head_d;
	uint64_t rot_a, rot_b;       // byte m-1: (tprime + 32 (m-1)) % 31 and % 33
};

inline HashK make_hashk(uint32_t k)
{
	HashK K;
	K.k = k;
	K.nblk = (k + 31) / 32;
	K.tprime = k - 32 * (K.nblk - 1);
	K.head_ra = (31 - (32 - K.tprime) % 31) % 31;
	K.head_rb = (33 - (32 - K.tprime) % 33) % 33;
	K.head_c = K.head_d = 0;
	for (uint32_t i = K.tprime; i < 32; i++) {
		K.head_c ^= srol_n(seed_of(0), 31 - i);
		K.head_d ^= srol_n(seed_of(3), i);
	}
	K.rot_a = K.rot_b = 0;
	for (uint32_t m = 1; m < K.nblk && m <= 8; m++) {
		K.rot_a |= (uint64_t)((K.tprime + 32 * (m - 1)) % 31) << (8 * (m - 1));
		K.rot_b |= (uint64_t)((K.tprime + 32 * (m - 1)) % 33) << (8 * (m - 1));
	}
	return K;
}

#if defined(__CUDACC__)
struct HitLoadV { // the one or two 16-byte groups of a record that hold three consecutive packed words, in flight
	uint4 g0, g1;
};

// rec: the record (16-byte aligned, ngroups uint4 long); w: index of the first of the three words inside the record (>= 1:
// word 0 is the length).  One 128-bit load when the three words lie in one group, two otherwise -- against three 32-bit
// loads, each its own L1 wavefront per lane.
__device__ __forceinline__ HitLoadV hit_issue_v(const uint4* __restrict__ rec, uint32_t ngroups, uint32_t w)
{
	HitLoadV h;
	const uint32_t g = w >> 2;
	h.g0 = __ldg(rec + min(g, ngroups - 1u));
	h.g1 = make_uint4(0u, 0u, 0u, 0u);
	if ((w & 3u) >= 2u && g + 1u < ngroups)
		h.g1 = __ldg(rec + g + 1u);
	return h;
}

__device__ __forceinline__ HitLoad hit_select(const HitLoadV& h, uint32_t w)
{
	const bool odd = (w & 1u) != 0, hi = (w & 2u) != 0;
	const uint32_t a0 = odd ? h.g0.y : h.g0.x, a1 = odd ? h.g0.z : h.g0.y, a2 = odd ? h.g0.w : h.g0.z, a3 = odd ? h.g1.x : h.g0.w,
	               a4 = odd ? h.g1.y : h.g1.x;
	HitLoad r;
	r.x0 = hi ? a2 : a0;
	r.x1 = hi ? a3 : a1;
	r.x2 = hi ? a4 : a2;
	return r;
}

// Canonical hash of the k-mer starting at base p of the record and ntComp (ntcard.cpp:132-145): the counter index inside the
// k's [2][2^rBits] sub-sketch, or kVoid when ntComp does not sample it.  h0: hit_issue_v(rec, ngroups, 1 + (p >> 4)).
__device__ __forceinline__ uint32_t hash_kmer_v(const HashK& K, const uint4* __restrict__ tab, uint32_t rBits, uint32_t S, const HitLoadV& h0,
    const uint4* __restrict__ rec, uint32_t ngroups, uint32_t p)
{
	uint32_t hh, hl;
	const uint32_t sh = (p & 15u) * 2u;
	const HitLoad x = hit_select(h0, 1u + (p >> 4));
	uint32_t w0 = __funnelshift_r(x.x0, x.x1, sh), w1 = __funnelshift_r(x.x1, x.x2, sh);
	uint32_t f0, f1, r0, r1;
	if (K.nblk == 1 && K.tprime == 32) { // k = 32: pure 32-bit path
		block_tables(tab, w0, w1, f0, f1, r0, r1);
		const bool rlt = r1 < f1 || (r1 == f1 && r0 < f0);
		hh = rlt ? r1 : f1;
		hl = rlt ? r0 : f0;
	} else {
		const uint32_t t = K.tprime;
		if (t <= 16) {
			w0 &= t == 16 ? 0xFFFFFFFFu : ((1u << (2u * t)) - 1u);
			w1 = 0;
		} else if (t < 32) {
			w1 &= (1u << (2u * (t - 16u))) - 1u;
		}
		block_tables(tab, w0, w1, f0, f1, r0, r1);
		uint64_t fh = (((uint64_t)f1 << 32) | f0) ^ K.head_c, rh = (((uint64_t)r1 << 32) | r0) ^ K.head_d;
		if (t < 32)
			fh = srol_ab(fh, K.head_ra, K.head_rb);
		for (uint32_t m = 1; m < K.nblk; m++) {
			const uint32_t o = p + t + 32u * (m - 1u), wm = 1u + (o >> 4), shm = (o & 15u) * 2u;
			const HitLoadV hv = hit_issue_v(rec, ngroups, wm);
			const HitLoad y = hit_select(hv, wm);
			block_tables(tab, __funnelshift_r(y.x0, y.x1, shm), __funnelshift_r(y.x1, y.x2, shm), f0, f1, r0, r1);
			const uint64_t FB = ((uint64_t)f1 << 32) | f0, RB = ((uint64_t)r1 << 32) | r0;
			fh = srol_ab(fh, 1, 32) ^ FB;
			const uint32_t ra = (uint32_t)(K.rot_a >> (8u * (m - 1u))) & 0xFFu, rb = (uint32_t)(K.rot_b >> (8u * (m - 1u))) & 0xFFu;
			rh ^= (ra | rb) ? srol_ab(RB, ra, rb) : RB;
		}
		const uint64_t hm = rh < fh ? rh : fh;
		hh = (uint32_t)(hm >> 32);
		hl = (uint32_t)hm;
	}
	const bool t0 = (hh >> (31 - S)) == 1u;
	const bool t1 = (hh >> (32 - S)) == ((1u << (S - 1)) - 1u);
	if (!(t0 || t1))
		return kVoid;
	return ((t1 ? 1u : 0u) << rBits) | (hl & ((1u << rBits) - 1u));
}
#endif

} // namespace pl
} // namespace ntc
