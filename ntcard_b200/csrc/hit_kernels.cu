// hit_kernels.cu -- stages H, F and A of the sketch pipeline (pipeline.h).
//
//   hit_kernel       reads the mask words written by the scan kernel, compacts the set bits into a queue,
//                    re-hashes every candidate k-mer in full 64 bits from the packed bases (byte-indexed tables
//                    in shared memory: 8 x 128-bit lookups per 32 bases), applies ntComp (ntcard.cpp:132-145)
//                    and appends the counter index to the binned hit log (block-sorted in shared memory so that
//                    the appends are coalesced runs).
//   fallback_kernel  tiles flagged by the scan kernel (records of different lengths): 64-bit recurrence, direct RED.
//   apply_kernel     the flush: slice by slice, stage the slice in L2 (zeros on the first flush after a reset,
//                    otherwise an L2 prefetch), then RED.ADD its log entries.  Pipelined over slices with a
//                    per-slice arrival counter, so no grid-wide barrier is needed.
#include <cuda_runtime.h>

#include "nthash_device.cuh"
#include "pipeline.h"
#include "sketch_common.cuh"

namespace ntc {
namespace pl {

namespace {

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

struct HitLoad { // a candidate with the packed words of its first 32-base block in flight
	uint32_t p, nwords;
	const uint32_t* __restrict__ b;
	uint32_t x0, x1, x2;
};

__device__ __forceinline__ HitLoad hit_issue(const HashCtx& c, uint32_t rec, uint32_t p, uint32_t nwords)
{
	HitLoad h;
	h.p = p;
	h.nwords = nwords;
	h.b = c.words + (uint64_t)rec * c.stride + 1;
	const uint32_t wi = (p + (c.k & 31u)) >> 4;
	h.x0 = __ldg(h.b + min(wi, nwords - 1));
	h.x1 = __ldg(h.b + min(wi + 1, nwords - 1));
	h.x2 = __ldg(h.b + min(wi + 2, nwords - 1));
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
__device__ __forceinline__ uint32_t hit_finish(const HashCtx& c, const HitLoad& h)
{
	const uint32_t t = c.k & 31u, M = c.k >> 5;
	uint32_t hh, hl; // canonical hash, high / low word
	if (t == 0 && M == 1) { // k = 32: pure 32-bit path
		const uint32_t sh = (h.p & 15u) * 2u;
		uint32_t f0, f1, r0, r1;
		block_tables(c.tab, __funnelshift_r(h.x0, h.x1, sh), __funnelshift_r(h.x1, h.x2, sh), f0, f1, r0, r1);
		const bool rlt = r1 < f1 || (r1 == f1 && r0 < f0);
		hh = rlt ? r1 : f1;
		hl = rlt ? r0 : f0;
	} else {
		uint64_t fh = 0, rh = 0;
		for (uint32_t i = 0; i < t; i++) {
			const uint32_t code = base_at(h.b, h.p + i);
			fh = srol(fh) ^ seed_of(code);
			rh ^= srol_n(seed_of(3u - code), i);
		}
		uint32_t x0 = h.x0, x1 = h.x1, x2 = h.x2;
		for (uint32_t m = 0; m < M; m++) {
			const uint32_t o = h.p + t + 32u * m, sh = (o & 15u) * 2u;
			if (m) {
				const uint32_t wi = o >> 4;
				x0 = __ldg(h.b + wi);
				x1 = __ldg(h.b + wi + 1);
				x2 = __ldg(h.b + min(wi + 2, h.nwords - 1));
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

// ------------------------------------------------------------------------------------------------
// hit kernel
// ------------------------------------------------------------------------------------------------
// A CTA is kHitGroups independent groups of 256 threads that share only the byte tables; each group works on its
// own unit (kHitRows... `rows_per_unit` consecutive mask rows = 1024 candidates on average) and synchronises with
// a named barrier, so the groups of an SM are in different phases and hide each other's latencies.
struct GroupSmem {
	uint32_t queue[kQueueCap];
	uint16_t rank[kQueueCap];
	uint32_t cnt[kMaxBins];
	uint32_t cur[kMaxBins], fill[kMaxBins], curpos[kMaxBins];                      // open block of every bin (persists over rounds)
	uint32_t cur0[kMaxBins], fill0[kMaxBins], first[kMaxBins], nvalid[kMaxBins];   // this round's placement
	uint32_t qn, ovf;
};
struct HitSmem {
	uint4 tab[8 * 256];
	GroupSmem g[kHitGroups];
};

constexpr int kHitBatch = 2; // candidates per thread whose loads are in flight together

__device__ __forceinline__ void group_sync(uint32_t g)
{
	asm volatile("bar.sync %0, %1;" ::"r"(g + 1u), "n"(kGroupThreads) : "memory");
}

// queue entry: slot (5 bits) | lane (5) << 5 | mask row inside the unit << 10
__device__ __forceinline__ void process_queue(GroupSmem& sm, const HitArgs& a, const HashCtx& c, uint32_t n, uint32_t row0, uint32_t g, uint32_t gtid)
{
	const Pool& P = a.pool;
	if (gtid < kMaxBins)
		sm.cnt[gtid] = 0;
	group_sync(g);
	// ---- pass 1: candidates -> counter indices, rank inside the bin --------------------------------------------
	for (uint32_t base = 0; base < n; base += kGroupThreads * kHitBatch) {
		HitLoad h[kHitBatch];
#pragma unroll
		for (int u = 0; u < kHitBatch; u++) {
			const uint32_t i = base + u * kGroupThreads + gtid;
			if (i < n) {
				const uint32_t e = sm.queue[i];
				const uint32_t s = e & 31u, ln = (e >> 5) & 31u, grow = row0 + (e >> 10);
				const uint32_t tile = grow / a.npos_max, p = grow - tile * a.npos_max;
				h[u] = hit_issue(c, tile * kTileRecs + s * 32u + ln, p, a.stride - 1u);
			}
		}
#pragma unroll
		for (int u = 0; u < kHitBatch; u++) {
			const uint32_t i = base + u * kGroupThreads + gtid;
			if (i < n) {
				const uint32_t idx = hit_finish(c, h[u]);
				sm.queue[i] = idx;
				if (idx != kVoid)
					sm.rank[i] = (uint16_t)atomicAdd(&sm.cnt[idx >> P.bin_shift], 1u);
			}
		}
	}
	group_sync(g);
	// ---- placement: pool blocks for what does not fit the open blocks; block lists of the slices ----------------
	if (gtid < P.nbins) {
		const uint32_t b = gtid;
		const uint32_t cb = sm.cnt[b];
		const uint32_t f0 = sm.fill[b], cu = sm.cur[b];
		const uint32_t space = kBlkEntries - f0; // an absent block has fill = kBlkEntries
		sm.fill0[b] = f0;
		sm.cur0[b] = cu;
		sm.first[b] = 0;
		sm.nvalid[b] = 0;
		const uint32_t slice = a.ki * P.nbins + b;
		uint32_t* list = P.slice_blocks + (size_t)slice * P.slice_cap;
		if (cb > space) {
			const uint32_t rem = cb - space, need = (rem + kBlkEntries - 1) / kBlkEntries;
			const uint32_t fst = atomicAdd(P.ctl + CTL_NEXT, need);
			const uint32_t at = atomicAdd(P.slice_nblk + slice, need);
			const uint32_t last_fill = rem - (need - 1) * kBlkEntries;
			if (cu != kVoid)
				list[sm.curpos[b]] = (cu << 9) | kBlkEntries; // the open block gets filled up
			uint32_t nv = 0;
			for (uint32_t j = 0; j < need; j++) {
				const uint64_t blk = (uint64_t)fst + j;
				const bool ok = blk < P.n_blocks && (uint64_t)at + j < P.slice_cap;
				if ((uint64_t)at + j < P.slice_cap)
					list[at + j] = ok ? (((uint32_t)blk << 9) | (j == need - 1 ? last_fill : kBlkEntries)) : kVoid;
				if (ok)
					nv = j + 1;
			}
			sm.first[b] = fst;
			sm.nvalid[b] = nv;
			if (nv == need) {
				sm.cur[b] = fst + need - 1;
				sm.fill[b] = last_fill;
				sm.curpos[b] = at + need - 1;
			} else { // pool exhausted: what did not get a block is added to the sketch directly (see pipeline.h)
				sm.cur[b] = kVoid;
				sm.fill[b] = kBlkEntries;
			}
		} else if (cb) {
			sm.fill[b] = f0 + cb;
			list[sm.curpos[b]] = (cu << 9) | (f0 + cb);
		}
	}
	group_sync(g);
	// ---- pass 2: entries into the pool blocks -----------------------------------------------------------------------
	for (uint32_t i = gtid; i < n; i += kGroupThreads) {
		const uint32_t idx = sm.queue[i];
		if (idx == kVoid)
			continue;
		const uint32_t b = idx >> P.bin_shift;
		const uint32_t v = sm.fill0[b] + sm.rank[i];
		uint32_t blk, o;
		if (v < kBlkEntries) {
			blk = sm.cur0[b];
			o = v;
		} else {
			const uint32_t q = (v - kBlkEntries) / kBlkEntries;
			o = (v - kBlkEntries) % kBlkEntries;
			blk = q < sm.nvalid[b] ? sm.first[b] + q : kVoid;
		}
		if (blk != kVoid) {
			P.entries[(size_t)blk * kBlkEntries + o] = idx;
		} else {
			atomicAdd(a.ctr_k + idx, 1u); // RED: legal, the sketch is materialised whenever the pool can run out
			P.ctl[CTL_DIRECT] = 1u;       // statistics only
		}
	}
	group_sync(g);
}

__device__ __forceinline__ void emit_bits(GroupSmem& sm, uint32_t x, uint32_t w, uint32_t& at)
{
	while (x) {
		const uint32_t s = __ffs(x) - 1;
		x &= x - 1;
		sm.queue[at++] = s | (w << 5);
	}
}

__global__ void __launch_bounds__(kHitThreads, 2) hit_kernel(const HitArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	HitSmem& S = *reinterpret_cast<HitSmem*>(smem_raw);
	const uint32_t tid = threadIdx.x, g = tid / kGroupThreads, gtid = tid % kGroupThreads;
	GroupSmem& sm = S.g[g];
	for (uint32_t i = tid; i < 8 * 256; i += kHitThreads)
		S.tab[i] = a.d_tab[i];
	if (gtid < kMaxBins) {
		sm.cur[gtid] = kVoid;
		sm.fill[gtid] = kBlkEntries;
	}
	__syncthreads();
	HashCtx c;
	c.words = a.words;
	c.stride = a.stride;
	c.k = a.k;
	c.rBits = a.pool.rBits;
	c.sBits = a.sBits;
	c.tab = S.tab;
	c.rot_a = a.rot_a;
	c.rot_b = a.rot_b;
	const uint64_t total_rows = (uint64_t)a.n_tiles * a.npos_max;
	const uint32_t RU = a.rows_per_unit;
	const uint32_t n_units = (uint32_t)((total_rows + RU - 1) / RU);
	constexpr int kW = 8; // mask words per thread in flight
	for (uint32_t unit = blockIdx.x * kHitGroups + g; unit < n_units; unit += gridDim.x * kHitGroups) {
		const uint32_t row0 = unit * RU;
		const uint32_t nw = (uint32_t)min((uint64_t)RU, total_rows - row0) * 32u;
		const uint32_t* m = a.masks + (size_t)row0 * 32u;
		if (gtid == 0) {
			sm.qn = 0;
			sm.ovf = 0;
		}
		group_sync(g);
		for (uint32_t w0 = 0; w0 < nw; w0 += kGroupThreads * kW) {
			uint32_t x[kW];
			uint32_t cnt = 0;
#pragma unroll
			for (int u = 0; u < kW; u++) {
				const uint32_t w = w0 + u * kGroupThreads + gtid;
				x[u] = w < nw ? __ldcs(m + w) : 0u;
			}
#pragma unroll
			for (int u = 0; u < kW; u++)
				cnt += __popc(x[u]);
			if (cnt) {
				uint32_t at = atomicAdd(&sm.qn, cnt);
				if (at + cnt <= kQueueCap) {
#pragma unroll
					for (int u = 0; u < kW; u++)
						emit_bits(sm, x[u], w0 + u * kGroupThreads + gtid, at);
				} else {
					sm.ovf = 1;
				}
			}
		}
		group_sync(g);
		if (!sm.ovf) {
			const uint32_t n = sm.qn;
			if (n)
				process_queue(sm, a, c, n, row0, g, gtid);
		} else {
			// skewed data: more candidates than the queue holds -> two mask rows (<= 2048 candidates) per round
			for (uint32_t w0 = 0; w0 < nw; w0 += 64) {
				group_sync(g);
				if (gtid == 0)
					sm.qn = 0;
				group_sync(g);
				const uint32_t w = w0 + gtid;
				const uint32_t x = (gtid < 64 && w < nw) ? __ldg(m + w) : 0u;
				if (x) {
					uint32_t at = atomicAdd(&sm.qn, (uint32_t)__popc(x));
					emit_bits(sm, x, w, at);
				}
				group_sync(g);
				const uint32_t n = sm.qn;
				if (n)
					process_queue(sm, a, c, n, row0, g, gtid);
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// fallback kernel: tiles the scan kernel flagged (records of different lengths)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fallback_kernel(const uint32_t* __restrict__ words, uint32_t stride, uint32_t n_rec, uint32_t n_tiles,
    const uint32_t* __restrict__ tile_info, const DevParams* __restrict__ P, uint32_t ki, uint32_t* __restrict__ ctr_k,
    const uint32_t* __restrict__ ctl)
{
	if (ctl[CTL_NFLAG] == 0)
		return;
	__shared__ KTab T;
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(&P->tab[ki]);
		uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
		for (uint32_t i = threadIdx.x; i < sizeof(KTab) / 4; i += blockDim.x)
			dst[i] = src[i];
	}
	__syncthreads();
	const uint32_t k = P->k[ki], rBits = P->rBits, sBits = P->sBits;
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		if (tile_info[tile] != kTileFlag)
			continue;
		for (uint32_t r = threadIdx.x; r < kTileRecs; r += blockDim.x) {
			const uint32_t rec = tile * kTileRecs + r;
			if (rec < n_rec) {
				const uint32_t* p = words + (uint64_t)rec * stride;
				process_piece_k(p + 1, __ldg(p), 0, k, T, ctr_k, rBits, sBits); // F1 was counted by the scan kernel
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// apply kernel (the flush)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void role_sync(uint32_t role)
{
	asm volatile("bar.sync %0, %1;" ::"r"(role + 1u), "n"(kApplyThreads / 2) : "memory");
}

// Two roles per CTA, decoupled through per-slice arrival counters in global memory:
//   stagers  (threads 0..255)   bring slice s into L2 -- write zeros (first flush after a reset) or issue L2
//                               prefetches -- at most kApplyAhead slices ahead of the slowest applier;
//   appliers (threads 256..511) RED.ADD the log entries of slice s once every CTA has staged it.
// In steady state the stagers stream the sketch at HBM write speed while the appliers' atomics hit L2.
constexpr uint32_t kApplyAhead = 2;

__global__ void __launch_bounds__(kApplyThreads) apply_kernel(const Pool P, uint32_t* __restrict__ counters, int force, uint32_t reserve_blocks)
{
	// ---- decide (every CTA reads the same values: they were written by earlier kernels) ----
	const uint32_t next = P.ctl[CTL_NEXT];
	const uint32_t used = min(next, P.n_blocks);
	const uint32_t state = P.ctl[CTL_STATE];
	if (!force) {
		const unsigned long long need = *P.cand / kBlkEntries + reserve_blocks;
		const bool tight = used + need > P.n_blocks;
		const bool flush = state == 0 ? (tight || P.ctl[CTL_NFLAG] != 0) : (tight && used > 0);
		if (!flush)
			return;
	} else if (state == 1 && next == 0) {
		return;
	}
	const bool zero_mode = state == 0;
	const uint32_t role = threadIdx.x / (kApplyThreads / 2), rtid = threadIdx.x % (kApplyThreads / 2), nrt = kApplyThreads / 2;
	const size_t slice_len = (size_t)1 << P.bin_shift; // counters per slice
	const size_t per_k = (size_t)2 << P.rBits;
	const uint32_t sparse_below = (uint32_t)(slice_len / 64 / kBlkEntries); // fewer entries than 1/8 of the slice's sectors: no prefetch

	if (role == 0) {
		// ================= stagers =================
		const size_t n16 = slice_len / 4; // 16-byte units per slice
		const size_t a0 = n16 * blockIdx.x / gridDim.x, a1 = n16 * (blockIdx.x + 1) / gridDim.x;
		for (uint32_t s = 0; s < P.n_slices; s++) {
			if (s >= kApplyAhead) { // L2 footprint: do not run more than kApplyAhead slices ahead of the appliers
				if (rtid == 0)
					while (ld_acquire(P.apply_done + s - kApplyAhead) < gridDim.x)
						__nanosleep(32);
				role_sync(role);
			}
			uint4* p = reinterpret_cast<uint4*>(counters + (size_t)(s / P.nbins) * per_k + (size_t)(s % P.nbins) * slice_len);
			if (zero_mode) {
				for (size_t i = a0 + rtid; i < a1; i += nrt)
					p[i] = make_uint4(0u, 0u, 0u, 0u);
				role_sync(role);
				if (rtid == 0) {
					__threadfence();
					atomicAdd(P.zero_done + s, 1u);
				}
			} else if (min(P.slice_nblk[s], P.slice_cap) >= sparse_below) {
				for (size_t i = a0 + (size_t)rtid * 256; i < a1; i += (size_t)nrt * 256) { // 4 KB per request
					const uint32_t bytes = (uint32_t)(min(a1 - i, (size_t)256) * 16);
					asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + i), "r"(bytes) : "memory");
				}
			}
		}
	} else {
		// ================= appliers =================
		const uint32_t lane = rtid & 31u, warp = rtid >> 5, nwarp = nrt / 32;
		for (uint32_t s = 0; s < P.n_slices; s++) {
			const uint32_t nb = min(P.slice_nblk[s], P.slice_cap);
			if (nb) {
				if (zero_mode) {
					if (rtid == 0)
						while (ld_acquire(P.zero_done + s) < gridDim.x)
							__nanosleep(32);
					role_sync(role);
				}
				uint32_t* ctr = counters + (size_t)(s / P.nbins) * per_k;
				const uint32_t* list = P.slice_blocks + (size_t)s * P.slice_cap;
				for (uint32_t j = blockIdx.x * nwarp + warp; j < nb; j += gridDim.x * nwarp) {
					const uint32_t le = __ldcs(list + j);
					if (le == kVoid)
						continue;
					const uint32_t f = min(le & 511u, kBlkEntries);
					const uint32_t* e = P.entries + (size_t)(le >> 9) * kBlkEntries;
					uint32_t v[kBlkEntries / 32];
#pragma unroll
					for (int u = 0; u < (int)(kBlkEntries / 32); u++)
						v[u] = lane + 32u * u < f ? __ldcs(e + lane + 32u * u) : kVoid;
#pragma unroll
					for (int u = 0; u < (int)(kBlkEntries / 32); u++)
						if (v[u] != kVoid)
							atomicAdd(ctr + v[u], 1u); // RED.ADD, L2 resident
				}
			}
			role_sync(role);
			if (rtid == 0)
				atomicAdd(P.apply_done + s, 1u);
		}
	}
	// ---- the last CTA to finish resets the pool ----
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		const uint32_t t = atomicAdd(P.ctl + CTL_TICKET, 1u);
		if (t == gridDim.x - 1) {
			for (uint32_t s = 0; s < P.n_slices; s++) {
				P.slice_nblk[s] = 0;
				P.zero_done[s] = 0;
				P.apply_done[s] = 0;
			}
			P.ctl[CTL_NEXT] = 0;
			P.ctl[CTL_STATE] = 1;
			P.ctl[CTL_TICKET] = 0;
			P.ctl[CTL_FLUSHES] += 1;
			__threadfence();
		}
	}
}

} // namespace

size_t hit_smem_bytes() { return sizeof(HitSmem); }

cudaError_t launch_hit(const HitArgs& a)
{
	static bool attr_set = false;
	if (!attr_set) {
		cudaError_t e = cudaFuncSetAttribute(hit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HitSmem));
		if (e != cudaSuccess)
			return e;
		attr_set = true;
	}
	hit_kernel<<<a.grid, kHitThreads, sizeof(HitSmem), a.stream>>>(a);
	return cudaGetLastError();
}

cudaError_t launch_fallback(const uint32_t* words, uint32_t stride, uint32_t n_rec, uint32_t n_tiles, const uint32_t* tile_info,
    const DevParams* d_params, uint32_t ki, uint32_t* ctr_k, const uint32_t* ctl, int n_sm, cudaStream_t st)
{
	const unsigned grid = n_tiles < (unsigned)n_sm * 4u ? n_tiles : (unsigned)n_sm * 4u;
	fallback_kernel<<<grid ? grid : 1, 256, 0, st>>>(words, stride, n_rec, n_tiles, tile_info, d_params, ki, ctr_k, ctl);
	return cudaGetLastError();
}

int apply_max_grid(int n_sm)
{
	int occ = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, apply_kernel, (int)kApplyThreads, 0) != cudaSuccess || occ < 1)
		occ = 1;
	return n_sm * (occ > 2 ? 2 : occ);
}

cudaError_t launch_apply(const Pool& pool, uint32_t* counters, int force, uint32_t reserve_blocks, unsigned grid, cudaStream_t st)
{
	apply_kernel<<<grid, kApplyThreads, 0, st>>>(pool, counters, force, reserve_blocks);
	return cudaGetLastError();
}

} // namespace pl
} // namespace ntc
