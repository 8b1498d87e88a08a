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
#include "hit_hash.cuh"
#include "launch.h"
#include "pipeline.h"
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

// ------------------------------------------------------------------------------------------------
// fast path: bit-sliced pre-filter (scan_kernel, MODE 2: T from the smallest register on the device; MODE 1: fixed T) + this kernel
// ------------------------------------------------------------------------------------------------
// Once every register is >= T - 1, only k-mers whose canonical hash has its top T bits zero can raise one (ntComp raises a
// register only when clz > register, nthll.cpp:94-95).  The scan kernel marks exactly those (1 in 2^(T-1): T = 9 -> 0.4 %) at
// ~5 instructions per k-mer instead of the ~100 of the 64-bit recurrence above; this kernel walks the mask words, re-hashes
// the marked k-mers in full (hit_hash.cuh) and applies ntComp to the registers in global memory (64 KB: L2 resident).
// Registers stay bit-exact: the filter has no false negatives, and every candidate goes through the same ntComp.
// Candidates are rare (T = 9: one mask word in 8 has a bit), so a warp first COMPACTS them: every lane pushes the set bits of its
// mask words into the warp's queue in shared memory, and whenever 32 are waiting all 32 lanes hash one each (walking the bits
// lane by lane left 1 lane in 8 busy: 0.8 ms per 10 M reads instead of ~0.1).
constexpr uint32_t kHllQueue = 32 + 32 * 4; // a full batch + what four more mask rows may add before draining (more than that: hashed in place)

struct HllHitArgs {
	const uint32_t* words;
	uint32_t stride, n_rec, n_tiles, npos_max;
	const uint32_t* masks;
	const uint32_t* tile_info;
	const uint4* d_tab;
	pl::HashK K;
	uint32_t nBits;
	uint32_t* regs;
};

__device__ __forceinline__ void hll_hit_one(const HllHitArgs& a, const uint4* __restrict__ tab, uint32_t tile, uint32_t r, uint32_t ln, uint32_t s,
    uint64_t low_mask)
{
	const uint32_t rec = tile * pl::kTileRecs + s * 32u + ln;
	if (rec >= a.n_rec)
		return; // slots past the end of the batch
	const uint32_t* rw = a.words + (uint64_t)rec * a.stride + 1;
	if (r + a.K.k > __ldg(rw - 1))
		return; // past the end of a shorter record of a mixed-length tile (it ran on the padding)
	const uint32_t last = a.stride - 2u;
	const uint64_t h = pl::canonical_hash<false>(a.K, tab, pl::hit_issue<false>(rw, r, last), rw, r, last);
	hll_comp(h, a.regs, low_mask);
}

// One warp per 16 mask rows of a tile (grid-stride).  The 2 KB of a unit are fetched with four 16-byte loads per lane, all in flight
// before the first is looked at (the walk is latency-bound otherwise: 152 MB of masks per 10 M reads against ~2 M marked k-mers);
// lane l's load q covers row c0 + 4q + l/8, records 4*(l%8) .. +3 of that row.  Queue entry: row << 10 | record lane << 5 | slot.
__global__ void __launch_bounds__(256, 4) hll_hit_kernel(const HllHitArgs a)
{
	__shared__ uint4 tab[8 * 256];
	__shared__ uint32_t q_e[8][kHllQueue];
	for (uint32_t i = threadIdx.x; i < 8 * 256; i += blockDim.x)
		tab[i] = a.d_tab[i];
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint64_t low_mask = ((uint64_t)1 << a.nBits) - 1;
	uint32_t* q = q_e[warp];
	constexpr uint32_t kChunk = 16; // rows per unit of work: a tile is shared by several warps, for balance on small batches
	constexpr int kLoads = kChunk / 4;
	const uint32_t cpt = (a.npos_max + kChunk - 1) / kChunk, n_units = a.n_tiles * cpt;
	const uint32_t sub_row = lane >> 3, rl0 = (lane & 7u) * 4u;
	for (uint32_t u = blockIdx.x * 8 + warp; u < n_units; u += gridDim.x * 8) {
		const uint32_t tile = u / cpt, c0 = (u - tile * cpt) * kChunk;
		const uint32_t info = __ldg(a.tile_info + tile);
		const uint32_t npos = info == pl::kTileFlag ? 0u : min(info, a.npos_max); // rows past it hold nothing (or are another tile's)
		const uint4* m = reinterpret_cast<const uint4*>(a.masks + (size_t)tile * a.npos_max * 32u) + lane;
		uint4 x[kLoads];
#pragma unroll
		for (int g = 0; g < kLoads; g++) {
			const uint32_t r = c0 + 4u * g + sub_row;
			x[g] = r < npos ? __ldcs(m + (size_t)(r - sub_row) * 8u) : make_uint4(0, 0, 0, 0);
		}
		uint32_t n = 0; // queued candidates of this warp (warp-uniform)
#pragma unroll
		for (int g = 0; g < kLoads; g++) {
			const uint32_t r = c0 + 4u * g + sub_row;
			uint32_t y[4] = {x[g].x, x[g].y, x[g].z, x[g].w};
			const uint32_t cnt = (uint32_t)(__popc(y[0]) + __popc(y[1]) + __popc(y[2]) + __popc(y[3]));
			const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, cnt);
			if (tot == 0)
				continue;
			if (n + tot > kHllQueue) { // skewed data: more marked k-mers in four rows than the queue has room for
#pragma unroll
				for (int j = 0; j < 4; j++)
					while (y[j]) {
						const uint32_t s = 31u - (uint32_t)__clz(y[j]);
						y[j] ^= 1u << s;
						hll_hit_one(a, tab, tile, r, rl0 + j, s, low_mask);
					}
				continue;
			}
			// exclusive prefix of cnt over the lanes: where this lane's candidates go in the queue
			uint32_t at = cnt;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, at, d);
				if (lane >= (uint32_t)d)
					at += t;
			}
			at = n + at - cnt;
#pragma unroll
			for (int j = 0; j < 4; j++)
				while (y[j]) {
					const uint32_t s = 31u - (uint32_t)__clz(y[j]);
					y[j] ^= 1u << s;
					q[at++] = (r << 10) | ((rl0 + j) << 5) | s;
				}
			n += tot;
			__syncwarp();
			while (n >= 32) { // a full batch: every lane hashes one candidate (taken from the end of the queue)
				const uint32_t e = q[n - 32 + lane];
				hll_hit_one(a, tab, tile, e >> 10, (e >> 5) & 31u, e & 31u, low_mask);
				n -= 32;
			}
			__syncwarp();
		}
		if (lane < n) { // what is left of this unit
			const uint32_t e = q[lane];
			hll_hit_one(a, tab, tile, e >> 10, (e >> 5) & 31u, e & 31u, low_mask);
		}
		__syncwarp();
	}
}

cudaError_t launch_hll_hit(const uint32_t* d_words, uint32_t stride, uint32_t n_rec, uint32_t n_tiles, uint32_t npos_max, const uint32_t* d_masks,
    const uint32_t* d_tile_info, const uint4* d_tab, uint32_t k, uint32_t nBits, uint8_t* d_regs, int n_sm, cudaStream_t st)
{
	HllHitArgs a;
	a.words = d_words;
	a.stride = stride;
	a.n_rec = n_rec;
	a.n_tiles = n_tiles;
	a.npos_max = npos_max;
	a.masks = d_masks;
	a.tile_info = d_tile_info;
	a.d_tab = d_tab;
	a.K = pl::make_hashk(k);
	a.nBits = nBits;
	a.regs = reinterpret_cast<uint32_t*>(d_regs);
	const unsigned units = n_tiles * ((npos_max + 15u) / 16u);
	static int per_sm = 0; // resident CTAs per SM (registers bound it); the same for every device of the box
	if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hll_hit_kernel, 256, 0) != cudaSuccess || per_sm < 1))
		per_sm = 1;
	const unsigned resident = (unsigned)(n_sm * per_sm);
	const unsigned ctas = (units + 7) / 8 < resident ? (units + 7) / 8 : resident;
	hll_hit_kernel<<<ctas ? ctas : 1, 256, 0, st>>>(a);
	return cudaGetLastError();
}

// smallest register (what the pre-filter may assume), one CTA
__global__ void __launch_bounds__(1024) hll_min_kernel(const uint32_t* __restrict__ regs, uint32_t n_words, uint32_t n_regs, uint32_t* __restrict__ out)
{
	__shared__ uint32_t s_min;
	if (threadIdx.x == 0)
		s_min = 255u;
	__syncthreads();
	uint32_t m = 0xFFFFFFFFu; // per-byte minimum
	if ((n_words & 3u) == 0) { // 16-byte loads, four in flight per thread: the kernel sits between two chunks of a batch
		const uint4* r4 = reinterpret_cast<const uint4*>(regs);
		const uint32_t n4 = n_words >> 2;
		for (uint32_t i0 = 0; i0 < n4; i0 += 4 * blockDim.x) {
			uint4 x[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const uint32_t i = i0 + j * blockDim.x + threadIdx.x;
				x[j] = i < n4 ? __ldcg(r4 + i) : make_uint4(~0u, ~0u, ~0u, ~0u);
			}
#pragma unroll
			for (int j = 0; j < 4; j++)
				m = __vminu4(m, __vminu4(__vminu4(x[j].x, x[j].y), __vminu4(x[j].z, x[j].w)));
		}
	} else {
		for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x)
			m = __vminu4(m, __ldcg(regs + i));
	}
	uint32_t v = 255u;
	for (uint32_t j = 0; j < 4 && j < n_regs; j++)
		v = min(v, (m >> (8 * j)) & 0xFFu);
	v = __reduce_min_sync(0xFFFFFFFFu, v);
	if ((threadIdx.x & 31u) == 0)
		atomicMin(&s_min, v);
	__syncthreads();
	if (threadIdx.x == 0)
		*out = s_min;
}

cudaError_t launch_hll_min(const uint8_t* d_regs, uint32_t nBits, uint32_t* d_out, cudaStream_t st)
{
	const uint32_t n_regs = 1u << nBits, n_words = nBits >= 2 ? (1u << (nBits - 2)) : 1u;
	hll_min_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const uint32_t*>(d_regs), n_words, n_regs, d_out);
	return cudaGetLastError();
}

} // namespace ntc
