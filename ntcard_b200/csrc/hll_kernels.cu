// hll_kernels.cu -- nthll's HyperLogLog registers on sm_100a.
//
// Reference: nthll.cpp:92-97 (ntComp: register h & (nBuck-1) = max(itself, clz(h & ~(nBuck-1))) when the upper bits are not
// all zero), :99-104 (ntRead: every canonical ntHash of the sequence), :213-241 (per-thread registers merged by max).
//
// Unlike ntcard's sketch, EVERY k-mer touches a register, so there is nothing to filter: one thread per piece rolls the
// 64-bit NTF64/NTR64 recurrence (as roll64_kernel does) and the registers of the whole CTA live in SHARED memory --
// 2^16 one-byte registers are 64 KB, three CTAs per SM.  A register only ever rises, and it rises O(log n) times, so the
// common case is one shared-memory byte load per k-mer; the rare rise is a compare-and-swap of the 32-bit word holding
// the byte (per-byte max = __vmaxu4).  A CTA starts from the registers already in HBM and merges back only the words
// that rose.  nBits > 17 does not fit shared memory: the same update runs on the registers in global memory (L2 resident).
#include "launch.h"
#include "sketch_common.cuh"

namespace ntc {

constexpr uint32_t HLL_BLOCK = 512;
constexpr uint32_t HLL_SMEM_MAX_BITS = 17; // 128 KB of registers per CTA

// regs: 2^nBits one-byte registers, viewed as 32-bit words (byte j of word w = register 4w + j, little endian)
template <typename WordPtr>
__device__ __forceinline__ void hll_raise(WordPtr regs, uint32_t idx, uint32_t run0)
{
	auto* w = regs + (idx >> 2);
	const uint32_t v = run0 << ((idx & 3u) * 8u);
	uint32_t old = *reinterpret_cast<volatile uint32_t*>(w);
	for (;;) {
		const uint32_t nw = __vmaxu4(old, v);
		if (nw == old)
			break;
		const uint32_t prev = atomicCAS(w, old, nw);
		if (prev == old)
			break;
		old = prev;
	}
}

// ntComp, nthll.cpp:92-97
__device__ __forceinline__ void hll_comp(uint64_t h, uint32_t* regs, uint64_t low_mask)
{
	const uint64_t top = h & ~low_mask;
	if (top) {
		const uint32_t run0 = (uint32_t)__clzll((long long)top);
		const uint32_t idx = (uint32_t)(h & low_mask);
		if (run0 > reinterpret_cast<volatile uint8_t*>(regs)[idx])
			hll_raise(regs, idx, run0);
	}
}

// ntRead (nthll.cpp:99-104) over k-mer starts [a, a + PIECE_STARTS) of one record: the iterator's rolling update
// (ntHashIterator.hpp:85 -> NTC64, nthash.hpp:275-279); records hold valid bases only.
__device__ __forceinline__ uint32_t hll_piece(const uint32_t* __restrict__ b, uint32_t len, uint32_t a, uint32_t k, const KTab& T,
    uint32_t* regs, uint64_t low_mask)
{
	if (len < k)
		return 0;
	const uint32_t ns = len - k + 1;
	if (a >= ns)
		return 0;
	const uint32_t e = min(a + PIECE_STARTS, ns);
	uint64_t fh = 0, rh = 0;
	{
		BaseStream s;
		s.open(b, a);
		for (uint32_t i = 0; i < k; i++)
			fh = srol(fh) ^ seed_of(s.next());
		for (uint32_t i = k; i-- > 0;)
			rh = srol(rh) ^ seed_of(3u - base_at(b, a + i));
	}
	hll_comp(rh < fh ? rh : fh, regs, low_mask);
	BaseStream so, si;
	so.open(b, a);
	if (a + 1 < e)
		si.open(b, a + k);
	for (uint32_t j = a + 1; j < e; j++) {
		const uint32_t idx = si.next() | (so.next() << 2);
		fh = srol(fh) ^ T.xf[idx];
		rh = sror(rh ^ T.xr[idx]);
		hll_comp(rh < fh ? rh : fh, regs, low_mask);
	}
	return e - a;
}

template <bool kRecordIsPiece, bool kShared>
__global__ void __launch_bounds__(HLL_BLOCK) hll_kernel(const uint32_t* __restrict__ words, const uint32_t* __restrict__ off, uint32_t stride,
    uint32_t n_rec, const uint32_t* __restrict__ piece_first, const uint32_t* __restrict__ piece_rec, const DevParams* __restrict__ P,
    uint32_t nBits, uint32_t* g_regs, unsigned long long* __restrict__ f1)
{
	extern __shared__ uint32_t s_regs[];
	__shared__ KTab tab;
	__shared__ uint32_t s_k;
	__shared__ unsigned long long s_f1;
	const uint32_t n_reg_words = nBits >= 2 ? (1u << (nBits - 2)) : 1u;
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(&P->tab[0]);
		uint32_t* dst = reinterpret_cast<uint32_t*>(&tab);
		for (uint32_t i = threadIdx.x; i < sizeof(KTab) / 4; i += blockDim.x)
			dst[i] = src[i];
		if (threadIdx.x == 0) {
			s_k = P->k[0];
			s_f1 = 0;
		}
		if (kShared)
			for (uint32_t i = threadIdx.x; i < n_reg_words; i += blockDim.x)
				s_regs[i] = __ldcg(g_regs + i); // any earlier value is a valid start: registers only rise
	}
	__syncthreads();
	uint32_t* regs = kShared ? s_regs : g_regs;
	const uint64_t low_mask = ((uint64_t)1 << nBits) - 1;
	const uint32_t k = s_k;
	const uint32_t n_pieces = kRecordIsPiece ? n_rec : piece_first[n_rec];
	unsigned long long mine = 0;
	// Block-uniform trip count (the merge below synchronises the CTA).  A register rises with probability ~ 1/t at its t-th
	// k-mer, so a CTA that only sees its own share keeps rising 'grid' times longer than the register file as a whole: the
	// CTAs exchange their registers through HBM after rounds 1, 2, 4, 8, ... (a constant expected number of rises per phase);
	// measured on 10 M reads in one launch, 444 private register files without the exchange: 31.7 ms.
	uint32_t round = 0;
	for (uint32_t base = blockIdx.x * blockDim.x; base < n_pieces; base += gridDim.x * blockDim.x, round++) {
		const uint32_t piece = base + threadIdx.x;
		if (piece < n_pieces) {
			uint32_t rec, a;
			if (kRecordIsPiece) {
				rec = piece;
				a = 0;
			} else {
				rec = piece_rec[piece];
				a = (piece - piece_first[rec]) * PIECE_STARTS;
			}
			const uint64_t ro = off ? (uint64_t)__ldg(off + rec) : (uint64_t)rec * stride;
			const uint32_t* r = words + ro;
			const uint64_t cap = off ? (uint64_t)__ldg(off + rec + 1) - ro : stride; // a corrupt length word must not leave the record
			const uint32_t len = (uint32_t)min((uint64_t)__ldg(r), cap ? (cap - 1) * 16 : 0);
			mine += hll_piece(r + 1, len, a, k, tab, regs, low_mask);
		}
		const bool last = base + gridDim.x * blockDim.x >= n_pieces || base + gridDim.x * blockDim.x < base;
		if (kShared && (last || ((round + 1) & round) == 0)) {
			__syncthreads();
			for (uint32_t i = threadIdx.x; i < n_reg_words; i += blockDim.x) { // merge by max, nthll.cpp:234-239
				const uint32_t v = s_regs[i];
				uint32_t old = __ldcg(g_regs + i), nw;
				for (;;) {
					nw = __vmaxu4(old, v);
					if (nw == old)
						break;
					const uint32_t prev = atomicCAS(g_regs + i, old, nw);
					if (prev == old)
						break;
					old = prev;
				}
				if (!last)
					s_regs[i] = nw; // and carry on from what every CTA has seen so far
			}
			__syncthreads();
		}
		if (last)
			break;
	}
	if (mine)
		atomicAdd(&s_f1, mine);
	__syncthreads();
	if (threadIdx.x == 0 && s_f1)
		atomicAdd(f1, s_f1);
}

cudaError_t launch_hll(const BatchView& b, bool record_is_piece, uint64_t n_pieces_bound, const uint32_t* d_piece_first,
    const uint32_t* d_piece_rec, const DevParams* d_params, uint32_t nBits, uint8_t* d_regs, unsigned long long* d_f1, int n_sm,
    cudaStream_t st)
{
	const bool shared = nBits <= HLL_SMEM_MAX_BITS;
	const size_t smem = shared ? (nBits >= 2 ? ((size_t)1 << nBits) : 4) : 0;
	// persistent CTAs: every CTA pays one load + one merge of its register file, so as few as fill the SMs
	unsigned per_sm = !shared ? 4u : smem <= (64u << 10) ? 3u : 1u;
	const uint64_t n = record_is_piece ? b.n_rec : n_pieces_bound;
	uint64_t grid = (n + HLL_BLOCK - 1) / HLL_BLOCK;
	grid = grid < 1 ? 1 : grid > (uint64_t)n_sm * per_sm ? (uint64_t)n_sm * per_sm : grid;
	uint32_t* regs = reinterpret_cast<uint32_t*>(d_regs);
#define NTC_HLL_LAUNCH(RP, SH)                                                                                                           \
	do {                                                                                                                                 \
		if (smem > (48u << 10)) {                                                                                                        \
			cudaError_t e = cudaFuncSetAttribute(hll_kernel<RP, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
			if (e != cudaSuccess)                                                                                                        \
				return e;                                                                                                                \
		}                                                                                                                                \
		hll_kernel<RP, SH><<<(unsigned)grid, HLL_BLOCK, smem, st>>>(b.words, b.off, b.stride, b.n_rec, d_piece_first, d_piece_rec, d_params, \
		    nBits, regs, d_f1);                                                                                                          \
	} while (0)
	if (record_is_piece) {
		if (shared)
			NTC_HLL_LAUNCH(true, true);
		else
			NTC_HLL_LAUNCH(true, false);
	} else {
		if (shared)
			NTC_HLL_LAUNCH(false, true);
		else
			NTC_HLL_LAUNCH(false, false);
	}
#undef NTC_HLL_LAUNCH
	return cudaGetLastError();
}

} // namespace ntc
