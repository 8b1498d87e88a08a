// bitslice_core.cuh -- the bit-sliced sampling filter (host/device, so it can be self-tested on the CPU).
//
// Idea.  ntComp (ntcard.cpp:132-145) looks only at the top s+1 bits of the canonical hash
// min(fh, rh), and those bits live in the UPPER 31-bit ring of the split-rotate hash
// (nthash.hpp:186-211: bits 63..33 and bits 32..0 rotate independently).  Because the top bits of a
// minimum are the minimum of the top bits, "is this k-mer sampled" needs only the upper rings of fh
// and rh.  The upper ring is GF(2)-linear in the per-base seeds, so we evaluate it BIT-SLICED: one
// 32-bit register per ring bit, bit s of every register belonging to "slot" s -- 32 different reads
// advance in lock step inside one thread.  A rotation of the ring then is a renaming of registers
// (free once the loop is unrolled by the ring length 31), and the update with the entering and the
// leaving base is ONE 3-input logic op per ring bit:
//
//     Uf'[r] = Uf[r-1] ^ SU[in][r]       ^ SU[out][r-k]            (NTF64, nthash.hpp:242-248)
//     Ur'[r] = Ur[r+1] ^ SU[3-in][r+1-k] ^ SU[3-out][r+1]          (NTR64, nthash.hpp:251-257)
//
// where SU[c][r] is bit 33+r of seed[c] and indices are mod 31.  Each SU[.][x] column is a boolean
// function of the base's two code bits (lo, hi); writing it as const ^ a1*lo ^ a2*hi ^ a4*(lo&hi),
// the non-constant part is one of 7 words prepared once per position (W[1..7]) and the constant is
// folded into the LOP3 truth table (XOR3 vs XNOR3).  So a position costs 62 LOP3 + 10 prep ops for
// 32 k-mers, instead of ~100 instructions PER k-mer for the 64-bit recurrence.
//
// Only k mod 31 enters the truth tables -> the scan is templated on KM = k % 31; everything else
// about k is runtime.  The k "virtual" positions before the first base (window not full yet) are
// run with leaving-base words = 0; the constants they wrongly inject are cancelled by the initial
// state (init_state()).
//
// Sampled k-mers (1 in 2^(s-1)) are then re-hashed in full 64 bits from the packed bases (hit path),
// which also decides table and bucket -- the filter only has to be free of false negatives.
#pragma once
#include <stdint.h>

#include "nthash_device.cuh"

namespace ntc {
namespace bs {

#if defined(__CUDACC__)
#define BS_HD __host__ __device__ __forceinline__
#else
#define BS_HD inline
#endif

constexpr uint64_t kSeed[4] = { 0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL, 0x20323ed082572324ULL, 0x295549f54be24456ULL };

// bit r (0..30) of the upper ring of seed[c]; r may be any integer (reduced mod 31)
constexpr int SU(int c, int r)
{
	return (int)((kSeed[c] >> (33 + ((r % 31) + 31) % 31)) & 1);
}
constexpr int mod31(int x) { return ((x % 31) + 31) % 31; }

// word selector of the non-constant part of c -> g(c), relative to g(0): bit0 = lo coefficient,
// bit1 = hi coefficient, bit2 = lo&hi coefficient
constexpr int sel(int g0, int g1, int g2, int g3)
{
	const int n1 = g1 ^ g0, n2 = g2 ^ g0, n3 = g3 ^ g0;
	return n1 | (n2 << 1) | ((n1 ^ n2 ^ n3) << 2);
}

// ---- forward strand, step q (TQ = q % 31), physical register j ------------------------------------
// after the step, physical j holds logical ring bit r = (j + q + 1) % 31
template <int KM, int TQ, int J> struct Fwd {
	static constexpr int r = mod31(J + TQ + 1);
	static constexpr int x = mod31(r - KM);                                          // leaving base reads SU[.][r-k]
	static constexpr int a = sel(SU(0, r), SU(1, r), SU(2, r), SU(3, r));           // entering-base word
	static constexpr int b = sel(SU(0, x), SU(1, x), SU(2, x), SU(3, x));           // leaving-base word
	static constexpr int c = SU(0, r) ^ SU(0, x);                                   // folded constant
};
// ---- reverse strand: after the step, physical j holds logical r = (j - q - 1) % 31 -----------------
template <int KM, int TQ, int J> struct Rev {
	static constexpr int r = mod31(J - TQ - 1);
	static constexpr int y = mod31(r + 1 - KM);                                      // entering base, complemented: SU[3-in][r+1-k]
	static constexpr int z = mod31(r + 1);                                           // leaving base, complemented: SU[3-out][r+1]
	static constexpr int a = sel(SU(3, y), SU(2, y), SU(1, y), SU(0, y));
	static constexpr int b = sel(SU(3, z), SU(2, z), SU(1, z), SU(0, z));
	static constexpr int c = SU(3, y) ^ SU(3, z);
};

struct Words { // W[i] for i = 0..7: 0, lo, hi, lo^hi, lh, lo^lh, hi^lh, lo^hi^lh
	uint32_t w[8];
};

BS_HD Words make_words(uint32_t lo, uint32_t hi)
{
	Words W;
	const uint32_t lh = lo & hi;
	W.w[0] = 0;
	W.w[1] = lo;
	W.w[2] = hi;
	W.w[3] = lo ^ hi;
	W.w[4] = lh;
	W.w[5] = lo ^ lh;
	W.w[6] = hi ^ lh;
	W.w[7] = lo ^ hi ^ lh;
	return W;
}

// state ^ a ^ b (^ 1 if C) as ONE logic op.  On the device this is an explicit lop3 so that the compiler
// cannot re-associate the XORs back into per-base expressions (it did: 2 LOP3 + stray NOTs per ring bit).
template <int C, int A, int B> BS_HD uint32_t xor3c(uint32_t s, const Words& in, const Words& out)
{
#if defined(__CUDA_ARCH__)
	uint32_t d;
	if (A != 0 && B != 0) {
		if (C)
			asm("lop3.b32 %0, %1, %2, %3, 0x69;" : "=r"(d) : "r"(s), "r"(in.w[A]), "r"(out.w[B]));
		else
			asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(s), "r"(in.w[A]), "r"(out.w[B]));
	} else if (A != 0 || B != 0) {
		const uint32_t w = A != 0 ? in.w[A] : out.w[B];
		if (C)
			asm("lop3.b32 %0, %1, %2, %2, 0xc3;" : "=r"(d) : "r"(s), "r"(w)); // ~(s ^ w)
		else
			asm("lop3.b32 %0, %1, %2, %2, 0x3c;" : "=r"(d) : "r"(s), "r"(w)); // s ^ w
	} else {
		d = C ? ~s : s;
	}
	return d;
#else
	const uint32_t v = s ^ in.w[A] ^ out.w[B];
	return C ? ~v : v;
#endif
}

struct State {
	uint32_t F[31];
	uint32_t R[31];
};

template <int KM, int TQ, int J> struct UpdateAll {
	static BS_HD void run(State& st, const Words& in, const Words& out)
	{
		using f = Fwd<KM, TQ, J>;
		using r = Rev<KM, TQ, J>;
		st.F[J] = xor3c<f::c, f::a, f::b>(st.F[J], in, out);
		st.R[J] = xor3c<r::c, r::a, r::b>(st.R[J], in, out);
		UpdateAll<KM, TQ, J + 1>::run(st, in, out);
	}
};
template <int KM, int TQ> struct UpdateAll<KM, TQ, 31> {
	static BS_HD void run(State&, const Words&, const Words&) {}
};

// One position: bases `in` enter, bases `out` leave (planes of 32 slots).
template <int KM, int TQ> BS_HD void step(State& st, uint32_t in_lo, uint32_t in_hi, uint32_t out_lo, uint32_t out_hi)
{
	const Words wi = make_words(in_lo, in_hi);
	const Words wo = make_words(out_lo, out_hi);
	UpdateAll<KM, TQ, 0>::run(st, wi, wo);
}

// ---- sampling test after step q: bit i (0 = MSB) of the upper rings --------------------------------
template <int TQ, int I> struct TopBit {
	static constexpr int jf = mod31(30 - I - (TQ + 1)); // physical register of forward logical bit 30-I
	static constexpr int jr = mod31(30 - I + (TQ + 1));
};

template <int TQ, int I, int S> struct Fold { // OR / AND chains over the top S bits
	static BS_HD void run(const State& st, uint32_t& nzA, uint32_t& nzB, uint32_t& eA, uint32_t& eB)
	{
		const uint32_t a = st.F[TopBit<TQ, I>::jf], b = st.R[TopBit<TQ, I>::jr];
		nzA |= a;
		nzB |= b;
		eA &= (I == 0) ? ~a : a;
		eB &= (I == 0) ? ~b : b;
		Fold<TQ, I + 1, S>::run(st, nzA, nzB, eA, eB);
	}
};
template <int TQ, int S> struct Fold<TQ, S, S> {
	static BS_HD void run(const State&, uint32_t&, uint32_t&, uint32_t&, uint32_t&) {}
};

// Slots whose k-mer (window ending at position q) is sampled by ntComp with sBits = S (S >= 2):
//   table 0: top S+1 bits of min(fh,rh) == 0..01      table 1: top S bits == 01..1
template <int TQ, int S> BS_HD uint32_t sampled_mask(const State& st)
{
	uint32_t nzA = 0, nzB = 0, eA = 0xFFFFFFFFu, eB = 0xFFFFFFFFu;
	Fold<TQ, 0, S>::run(st, nzA, nzB, eA, eB);
	const uint32_t aS = st.F[TopBit<TQ, S>::jf], bS = st.R[TopBit<TQ, S>::jr];   // bit S+1 from the top
	const uint32_t a0 = st.F[TopBit<TQ, 0>::jf], b0 = st.R[TopBit<TQ, 0>::jr];
	// min(A,B) == 1  <=>  (A==1 && B>=1) || (B==1 && A>=1), A = top S+1 bits
	const uint32_t t0 = (~nzA & aS & (nzB | bS)) | (~nzB & bS & (nzA | aS));
	// min(A',B') == 01..1 <=> (A'==v && (B' top bit || B'==v)) || (B'==v && A' top bit)
	const uint32_t t1 = (eA & (b0 | eB)) | (eB & a0);
	return t0 | t1;
}

// nthll pre-filter (nthll.cpp:92-97): a k-mer can only raise a HyperLogLog register if its leading-zero count exceeds the
// register, so once every register is >= T - 1 only hashes whose top T bits are all zero matter.  Top T <= 31 bits of
// min(fh, rh) are zero iff those of fh are or those of rh are.
template <int TQ, int I, int T> struct FoldOr {
	static BS_HD void run(const State& st, uint32_t& nzA, uint32_t& nzB)
	{
		nzA |= st.F[TopBit<TQ, I>::jf];
		nzB |= st.R[TopBit<TQ, I>::jr];
		FoldOr<TQ, I + 1, T>::run(st, nzA, nzB);
	}
};
template <int TQ, int T> struct FoldOr<TQ, T, T> {
	static BS_HD void run(const State&, uint32_t&, uint32_t&) {}
};
template <int TQ, int T> BS_HD uint32_t zero_top_mask(const State& st)
{
	uint32_t nzA = 0, nzB = 0;
	FoldOr<TQ, 0, T>::run(st, nzA, nzB);
	return ~nzA | ~nzB;
}

// The same with the number of bits chosen at run time (warp-uniform lev = 0..5: no filter, top 5, 7, 9, 11, 13 bits), so that one
// kernel serves every stage of a run and picks its filter from the smallest register as it stands on the device.
template <int TQ> BS_HD uint32_t zero_top_mask_lev(const State& st, uint32_t lev)
{
	uint32_t a = 0, b = 0, nzA, nzB;
	FoldOr<TQ, 0, 5>::run(st, a, b);
	nzA = lev >= 1 ? a : 0u, nzB = lev >= 1 ? b : 0u;
	a = b = 0;
	FoldOr<TQ, 5, 7>::run(st, a, b);
	nzA |= lev >= 2 ? a : 0u, nzB |= lev >= 2 ? b : 0u;
	a = b = 0;
	FoldOr<TQ, 7, 9>::run(st, a, b);
	nzA |= lev >= 3 ? a : 0u, nzB |= lev >= 3 ? b : 0u;
	a = b = 0;
	FoldOr<TQ, 9, 11>::run(st, a, b);
	nzA |= lev >= 4 ? a : 0u, nzB |= lev >= 4 ? b : 0u;
	a = b = 0;
	FoldOr<TQ, 11, 13>::run(st, a, b);
	nzA |= lev >= 5 ? a : 0u, nzB |= lev >= 5 ? b : 0u;
	return ~nzA | ~nzB;
}

// Initial state cancelling the constants the k virtual steps inject (see header comment):
//   forward : E0 = XOR_{i<k} sror^{1+i}( srol^k(seed[A]) )        reverse : E0 = XOR_{i<k} srol^i( seed[T] )
// Physical register j holds logical bit j at q = 0.  Returns broadcast words (0 / 0xFFFFFFFF).
inline void init_state(unsigned k, uint32_t F0[31], uint32_t R0[31])
{
	uint64_t ef = 0, er = 0;
	uint64_t v = srol_n(seed_of(0), k);
	uint64_t w = seed_of(3);
	for (unsigned i = 0; i < k; i++) {
		v = sror(v);
		ef ^= v;
		er ^= w;
		w = srol(w);
	}
	for (int j = 0; j < 31; j++) {
		F0[j] = ((ef >> (33 + j)) & 1) ? 0xFFFFFFFFu : 0u;
		R0[j] = ((er >> (33 + j)) & 1) ? 0xFFFFFFFFu : 0u;
	}
}

// ---- 32x32 bit-matrix transpose: out[o] bit s = in[s] bit o -----------------------------------------
// Five butterfly stages; the 16- and 8-bit ones are byte permutes (PRMT), the rest shift+select.
BS_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
	return __byte_perm(a, b, sel);
#else
	uint64_t v = ((uint64_t)b << 32) | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; i++)
		r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
	return r;
#endif
}

// m ? x : y, bit by bit, as ONE logic op with the mask in a register.  Written as (x & M) | (y & ~M) with a literal mask the
// compiler emits two LOP3 per output (an instruction takes one immediate): 96 extra ALU instructions per 16 positions.
BS_HD uint32_t bitsel(uint32_t x, uint32_t y, uint32_t m)
{
#if defined(__CUDA_ARCH__)
	uint32_t d;
	asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(x), "r"(y), "r"(m));
	return d;
#else
	return (x & m) | (y & ~m);
#endif
}

template <int JJ, uint32_t M> BS_HD void tr_stage(uint32_t (&A)[32])
{
	uint32_t m = M;
#if defined(__CUDA_ARCH__)
	asm volatile("" : "+r"(m)); // keep the mask in a register: the select below needs three register operands
#pragma unroll
#endif
	for (int s = 0; s < 32; s++) {
		if (s & JJ)
			continue;
		const uint32_t a = A[s], b = A[s + JJ];
#if defined(__CUDA_ARCH__)
		// the shifts as multiplications: IMAD / IMAD.HI issue on the FMA pipe, which is idle, instead of the ALU pipe, which bounds the kernel
		uint32_t bl, ar;
		asm("mul.lo.u32 %0, %1, %2;" : "=r"(bl) : "r"(b), "r"(1u << JJ));
		asm("mul.hi.u32 %0, %1, %2;" : "=r"(ar) : "r"(a), "r"(1u << (32 - JJ)));
		A[s] = bitsel(a, bl, m);
		A[s + JJ] = bitsel(ar, b, m);
#else
		A[s] = bitsel(a, b << JJ, m);
		A[s + JJ] = bitsel(a >> JJ, b, m);
#endif
	}
}

BS_HD void transpose32(uint32_t (&A)[32])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int s = 0; s < 16; s++) {
		const uint32_t a = A[s], b = A[s + 16];
		A[s] = byte_perm(a, b, 0x5410);
		A[s + 16] = byte_perm(a, b, 0x7632);
	}
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int s = 0; s < 32; s++) {
		if (s & 8)
			continue;
		const uint32_t a = A[s], b = A[s + 8];
		A[s] = byte_perm(a, b, 0x6240);
		A[s + 8] = byte_perm(a, b, 0x7351);
	}
	tr_stage<4, 0x0F0F0F0Fu>(A);
	tr_stage<2, 0x33333333u>(A);
	tr_stage<1, 0x55555555u>(A);
}

} // namespace bs
} // namespace ntc
