// hit_hash.cuh -- the full 64-bit canonical ntHash of one sampled k-mer from the packed bases, and ntComp
// (ntcard.cpp:132-145) on it: shared by the stand-alone hit kernel (hit_kernels.cu) and the fused sketch kernel
// (fused_kernel.cuh).  Byte-indexed tables in shared memory: 8 x 128-bit lookups per 32 bases.
//
// Block-wise for ANY k.  k = tprime + 32 (nblk - 1), 1 <= tprime <= 32: the first tprime bases form the HEAD block, put through
// the byte tables with the other 32 - tprime codes masked to 0 ('A'); what those A's contribute is a constant per k
// (head_c / head_d), and the head's hash sits 32 - tprime rotations too far left:
//   fh = sror^(32-tprime)( FB(head) ^ head_c ),  rh = RB(head) ^ head_d;   then per full block m = 1 .. nblk-1:
//   fh = srol^32(fh) ^ FB_m,                      rh ^= srol^(tprime + 32 (m-1)) RB_m
// with FB = XOR_i srol^(31-i) seed[c_i], RB = XOR_i srol^i seed[3-c_i] over the 32 codes of a block (NTF64 / NTR64 base forms,
// nthash.hpp:220-239, regrouped; checked for every k class by tests/test_parity_gpu.py::test_bitslice_every_k_class).  No per-base loop:
// k = 25 costs one block like k = 32, k = 31 one, k = 64 two.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nthash_device.cuh"
#include "pipeline.h"

namespace ntc {
namespace pl {

constexpr size_t kTabBytes = 8 * 256 * 16;

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
// srol applied n times, for n given as (a, b) = (n % 31, n % 33): rotate the upper ring by a, the lower by b.
__device__ __forceinline__ uint64_t srol_ab(uint64_t v, uint32_t a, uint32_t b)
{
	uint32_t hi = (uint32_t)(v >> 33);
	uint64_t lo = v & 0x1FFFFFFFFull;
	hi = ((hi << a) | (hi >> (31u - a))) & 0x7FFFFFFFu;
	lo = ((lo << b) | (lo >> (33u - b))) & 0x1FFFFFFFFull;
	return ((uint64_t)hi << 33) | lo;
}

struct HitLoad { // three consecutive packed words: 32 bases at any 2-bit offset
	uint32_t x0, x1, x2;
};

// 32 bases (two packed words) through the byte tables
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

// ntComp on the canonical hash (hh = high word, hl = low word): both tests look at the top S+1 <= 32 bits; the bucket at the low
// rBits <= 30 bits.  Returns the counter index inside the k's [2][2^rBits] sub-sketch, or kVoid when not sampled.
__device__ __forceinline__ uint32_t ntcomp_index(uint32_t hh, uint32_t hl, uint32_t rBits, uint32_t S)
{
	const bool t0 = (hh >> (31 - S)) == 1u;
	const bool t1 = (hh >> (32 - S)) == ((1u << (S - 1)) - 1u);
	if (!(t0 || t1))
		return kVoid;
	return ((t1 ? 1u : 0u) << rBits) | (hl & ((1u << rBits) - 1u)); // the later test wins (ntcard.cpp:136-139)
}

// The head block: x = the three words that hold bases p .. p+31 of the record.  For k = 32 this is the whole hash.
__device__ __forceinline__ void hash_head(const HashK& K, const uint4* __restrict__ tab, const HitLoad& x, uint32_t p, uint64_t& fh, uint64_t& rh)
{
	const uint32_t sh = (p & 15u) * 2u, t = K.tprime;
	uint32_t w0 = __funnelshift_r(x.x0, x.x1, sh), w1 = __funnelshift_r(x.x1, x.x2, sh);
	if (t <= 16) {
		w0 &= t == 16 ? 0xFFFFFFFFu : ((1u << (2u * t)) - 1u);
		w1 = 0;
	} else if (t < 32) {
		w1 &= (1u << (2u * (t - 16u))) - 1u;
	}
	uint32_t f0, f1, r0, r1;
	block_tables(tab, w0, w1, f0, f1, r0, r1);
	fh = (((uint64_t)f1 << 32) | f0) ^ K.head_c;
	rh = (((uint64_t)r1 << 32) | r0) ^ K.head_d;
	if (t < 32)
		fh = srol_ab(fh, K.head_ra, K.head_rb);
}

// One more full block: y = the three words that hold its 32 bases, o = index of its first base in the record, m >= 1
__device__ __forceinline__ void hash_block(const HashK& K, const uint4* __restrict__ tab, const HitLoad& y, uint32_t o, uint32_t m, uint64_t& fh,
    uint64_t& rh)
{
	const uint32_t sh = (o & 15u) * 2u;
	uint32_t f0, f1, r0, r1;
	block_tables(tab, __funnelshift_r(y.x0, y.x1, sh), __funnelshift_r(y.x1, y.x2, sh), f0, f1, r0, r1);
	const uint64_t FB = ((uint64_t)f1 << 32) | f0, RB = ((uint64_t)r1 << 32) | r0;
	fh = srol_ab(fh, 1, 32) ^ FB;
	const uint32_t ra = (uint32_t)(K.rot_a >> (8u * (m - 1u))) & 0xFFu, rb = (uint32_t)(K.rot_b >> (8u * (m - 1u))) & 0xFFu;
	rh ^= (ra | rb) ? srol_ab(RB, ra, rb) : RB;
}

// ---- 32-bit word loads (hit kernel: the tile staged in shared memory, or global memory) ---------------------------------------
template <bool kStaged> __device__ __forceinline__ uint32_t ld_word(const uint32_t* p)
{
	return kStaged ? *p : __ldg(p); // staged: the tile's packed reads sit in shared memory
}

// rec_words: the base words of the record (word 0 = length skipped); o: first base of the block; last: index of the last base
// word that may be read (the record's slot in the uniform-stride batch)
template <bool kStaged>
__device__ __forceinline__ HitLoad hit_issue(const uint32_t* __restrict__ rec_words, uint32_t o, uint32_t last)
{
	HitLoad h;
	const uint32_t wi = o >> 4;
	h.x0 = ld_word<kStaged>(rec_words + min(wi, last));
	h.x1 = ld_word<kStaged>(rec_words + min(wi + 1, last));
	h.x2 = ld_word<kStaged>(rec_words + min(wi + 2, last));
	return h;
}

// Canonical hash (NTC64, nthash.hpp:275-279: min of the two strands) of the k-mer starting at base p.
// h0 = hit_issue(rec_words, p, last), requested earlier.
template <bool kStaged>
__device__ __forceinline__ uint64_t canonical_hash(const HashK& K, const uint4* __restrict__ tab, const HitLoad& h0, const uint32_t* __restrict__ rec_words,
    uint32_t p, uint32_t last)
{
	uint64_t fh, rh;
	hash_head(K, tab, h0, p, fh, rh);
	for (uint32_t m = 1; m < K.nblk; m++) {
		const uint32_t o = p + K.tprime + 32u * (m - 1u);
		hash_block(K, tab, hit_issue<kStaged>(rec_words, o, last), o, m, fh, rh);
	}
	return rh < fh ? rh : fh;
}

// ... and ntComp on it
template <bool kStaged>
__device__ __forceinline__ uint32_t hit_finish(const HashK& K, const uint4* __restrict__ tab, uint32_t rBits, uint32_t S, const HitLoad& h0,
    const uint32_t* __restrict__ rec_words, uint32_t p, uint32_t last)
{
	const uint64_t hm = canonical_hash<kStaged>(K, tab, h0, rec_words, p, last);
	return ntcomp_index((uint32_t)(hm >> 32), (uint32_t)hm, rBits, S);
}

// ---- 128-bit gathers (fused kernel) ---------------------------------------------------------------------------------------------
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

// h0: hit_issue_v(rec, ngroups, 1 + (p >> 4)), requested earlier
__device__ __forceinline__ uint32_t hash_kmer_v(const HashK& K, const uint4* __restrict__ tab, uint32_t rBits, uint32_t S, const HitLoadV& h0,
    const uint4* __restrict__ rec, uint32_t ngroups, uint32_t p)
{
	uint64_t fh, rh;
	hash_head(K, tab, hit_select(h0, 1u + (p >> 4)), p, fh, rh);
	for (uint32_t m = 1; m < K.nblk; m++) {
		const uint32_t o = p + K.tprime + 32u * (m - 1u), wm = 1u + (o >> 4);
		hash_block(K, tab, hit_select(hit_issue_v(rec, ngroups, wm), wm), o, m, fh, rh);
	}
	const uint64_t hm = rh < fh ? rh : fh;
	return ntcomp_index((uint32_t)(hm >> 32), (uint32_t)hm, rBits, S);
}
#endif

} // namespace pl
} // namespace ntc
