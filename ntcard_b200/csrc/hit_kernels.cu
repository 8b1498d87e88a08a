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
#include <cstdlib>

#include "hit_hash.cuh"
#include "pipeline.h"
#include "sketch_common.cuh"

namespace ntc {
namespace pl {

namespace {

// ------------------------------------------------------------------------------------------------
// hit kernel
// ------------------------------------------------------------------------------------------------
// A CTA is kHitGroups independent groups of 256 threads that share only the byte tables; each group works on its
// own unit (a run of mask rows holding ~2048 candidates on average) and synchronises with a named barrier, so the
// groups of an SM are in different phases and hide each other's latencies.
//
// Log appends are optimistic: every bin of a group has an open pool block with at least kBlkSlack free entries at
// the start of a round, a hit takes its slot with one shared-memory atomic and is stored straight to the block.
// The rare hit that finds its block full waits in an overflow list for the end of the round, when thread b tops
// bin b up: a block with less than kBlkSlack free entries is closed and the spare block -- requested one round
// earlier, so its allocation latency is never waited for -- takes its place.
// shared memory of a group: [queue qcap][cur, fill, curpos: kMaxBins each][qn, ovf, n_ovf, pad][mbarrier 16 B]
// [staged tile: 1024 records x stride words (staged kernel only)]
struct GroupSmem {
	uint32_t* queue;
	uint32_t *cur, *fill, *curpos; // open block of every bin (persists over rounds)
	uint32_t* scal;                // [0] qn, [1] ovf, [2] n_ovf
	uint64_t* mbar;
	uint32_t* tile;
};
constexpr uint32_t kStagedQueueCap = 3072; // one tile: 1024 x ~119..145 positions / 64 = 1900..2320 candidates at s = 7
__host__ __device__ constexpr size_t group_smem_fixed(uint32_t qcap) { return ((size_t)qcap + 3 * kMaxBins + 4) * 4 + 16; }

__device__ __forceinline__ GroupSmem group_smem(unsigned char* base, uint32_t qcap)
{
	GroupSmem g;
	g.queue = reinterpret_cast<uint32_t*>(base);
	g.cur = g.queue + qcap;
	g.fill = g.cur + kMaxBins;
	g.curpos = g.fill + kMaxBins;
	g.scal = g.curpos + kMaxBins;
	g.mbar = reinterpret_cast<uint64_t*>(g.scal + 4);
	g.tile = reinterpret_cast<uint32_t*>(g.mbar + 2);
	return g;
}

constexpr uint32_t kPendingBit = 0x80000000u; // queue entry after its round: idx | kPendingBit = still to be appended

constexpr int kHitBatch = 4; // candidates per thread whose loads are in flight together

__device__ __forceinline__ void group_sync(uint32_t g)
{
	asm volatile("bar.sync %0, %1;" ::"r"(g + 1u), "n"(kGroupThreads) : "memory");
}

struct Unit {            // the mask rows a group works on
	uint32_t tile0;      // first tile
	uint32_t r0;         // first row inside tile0 (one-tile units), 0 otherwise
	uint32_t nrows;
	bool multi;          // rows run over several whole tiles (row r -> tile0 + r / npos_max)
};

struct Spare {           // thread b's spare block for bin b, in registers
	uint32_t blk, at;
	bool have;
};

// queue entry: slot (5 bits) | lane (5) << 5 | mask row inside the unit << 10
// rl: the record's index inside its tile
__device__ __forceinline__ void decode(const HitArgs& a, const Unit& U, uint32_t e, uint32_t& rec, uint32_t& rl, uint32_t& p)
{
	const uint32_t s = e & 31u, ln = (e >> 5) & 31u, r = e >> 10;
	uint32_t tile = U.tile0;
	p = U.r0 + r;
	if (U.multi) {
		const uint32_t t = r / a.npos_max;
		tile += t;
		p = r - t * a.npos_max;
	}
	rl = s * 32u + ln;
	rec = tile * kTileRecs + rl;
}

// one log entry: a slot in the bin's open block; false when the block is full (the entry is retried after the top-up)
__device__ __forceinline__ bool append(const GroupSmem& sm, const HitArgs& a, uint32_t idx)
{
	const Pool& P = a.pool;
	const uint32_t b = idx >> P.bin_shift;
	const uint32_t cu = sm.cur[b];
	if (cu == kVoid) { // the pool ran out: the sketch is materialised whenever that can happen (pipeline.h) -> RED
		atomicAdd(a.ctr_k + idx, 1u);
		P.ctl[CTL_DIRECT] = 1u; // statistics only
		return true;
	}
	const uint32_t slot = atomicAdd(&sm.fill[b], 1u);
	if (slot < kBlkEntries) {
		P.entries[(size_t)cu * kBlkEntries + slot] = idx;
		return true;
	}
	return false;
}

// end of a round: thread b closes bin b's block if it is (nearly) full and opens the spare
__device__ __forceinline__ void top_up(const GroupSmem& sm, const HitArgs& a, uint32_t b, Spare& sp)
{
	const Pool& P = a.pool;
	const uint32_t slice = a.ki * P.nbins + b;
	uint32_t* list = P.slice_blocks + (size_t)slice * P.slice_cap;
	const uint32_t cu = sm.cur[b];
	const uint32_t f = min(sm.fill[b], kBlkEntries);
	if (cu != kVoid) {
		list[sm.curpos[b]] = (cu << 9) | f;
		if (f + kBlkSlack <= kBlkEntries) {
			sm.fill[b] = f;
			return;
		}
	}
	// open the spare (a void spare = pool exhausted: the bin goes to direct increments)
	if (!sp.have) { // only at the very first round of the kernel
		sp.blk = atomicAdd(P.ctl + CTL_NEXT, 1u);
		sp.at = atomicAdd(P.slice_nblk + slice, 1u);
	}
	const bool ok = sp.blk < P.n_blocks && sp.at < P.slice_cap;
	if (sp.at < P.slice_cap)
		list[sp.at] = ok ? (sp.blk << 9) : kVoid;
	sm.cur[b] = ok ? sp.blk : kVoid;
	sm.curpos[b] = sp.at;
	sm.fill[b] = 0;
	// request the next spare now; nobody waits for the answer before the next top-up
	sp.blk = atomicAdd(P.ctl + CTL_NEXT, 1u);
	sp.at = atomicAdd(P.slice_nblk + slice, 1u);
	sp.have = true;
}

template <bool kStaged>
__device__ __forceinline__ void process_queue(const GroupSmem& sm, const HitArgs& a, const uint4* __restrict__ tab, uint32_t n, const Unit& U, uint32_t g,
    uint32_t gtid, Spare& sp)
{
	const Pool& P = a.pool;
	const uint32_t last = a.stride - 2u;
	// ---- candidates -> counter indices -> log ------------------------------------------------------------------
	for (uint32_t base = 0; base < n; base += kGroupThreads * kHitBatch) {
		HitLoad h[kHitBatch];
#pragma unroll
		for (int u = 0; u < kHitBatch; u++) {
			const uint32_t i = base + u * kGroupThreads + gtid;
			if (i < n) {
				uint32_t rec, rl, p;
				decode(a, U, sm.queue[i], rec, rl, p);
				const uint32_t* rw = kStaged ? sm.tile + rl * a.stride + 1 : a.words + (uint64_t)min(rec, a.n_rec - 1u) * a.stride + 1;
				h[u] = hit_issue<kStaged>(rw, p, last);
			}
		}
#pragma unroll
		for (int u = 0; u < kHitBatch; u++) {
			const uint32_t i = base + u * kGroupThreads + gtid;
			if (i < n) {
				uint32_t rec, rl, p;
				decode(a, U, sm.queue[i], rec, rl, p);
				const uint32_t* rw = kStaged ? sm.tile + rl * a.stride + 1 : a.words + (uint64_t)min(rec, a.n_rec - 1u) * a.stride + 1;
				// slots past the end of the batch, and positions past the end of a record of a mixed-length tile (scan_kernel.cuh)
				const uint32_t idx = (rec < a.n_rec && p + a.k <= (kStaged ? rw[-1] : __ldg(rw - 1))) ? hit_finish<kStaged>(a.hk, tab, P.rBits, a.sBits, h[u], rw, p, last) : kVoid;
				uint32_t after = 0;
				if (idx != kVoid && !append(sm, a, idx)) {
					after = idx | kPendingBit;
					atomicAdd(&sm.scal[2], 1u);
				}
				sm.queue[i] = after;
			}
		}
	}
	group_sync(g);
	// ---- top up the bins; place what overflowed (rare) ------------------------------------------------------------
	for (;;) {
		const uint32_t n_ovf = sm.scal[2];
		if (gtid < P.nbins)
			top_up(sm, a, gtid, sp);
		if (n_ovf == 0)
			break;
		group_sync(g);
		if (gtid == 0)
			sm.scal[2] = 0;
		group_sync(g);
		for (uint32_t i = gtid; i < n; i += kGroupThreads) {
			const uint32_t e = sm.queue[i];
			if (e & kPendingBit) {
				if (append(sm, a, e & ~kPendingBit))
					sm.queue[i] = 0;
				else
					atomicAdd(&sm.scal[2], 1u);
			}
		}
		group_sync(g);
	}
}

__device__ __forceinline__ void emit_bits(const GroupSmem& sm, uint32_t x, uint32_t w, uint32_t& at)
{
	const uint32_t hi = w << 5;
	while (x) {
		const uint32_t s = 31u - (uint32_t)__clz(x); // highest set bit: one FLO (the lowest needs a BREV first)
		x ^= 1u << s;
		sm.queue[at++] = s | hi;
	}
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// kStaged: a unit is one whole tile whose packed reads (1024 x stride words, <= 48 KB) are copied to shared memory by
// one bulk-async (TMA) request while the group enumerates the tile's mask words; the candidates' bases are then
// gathered from shared memory instead of with three uncoalesced global loads each (the L1 tag stage is the limit
// there: 32 different lines per warp instruction).
template <bool kStaged, int kGroups>
__global__ void __launch_bounds__(kGroups * kGroupThreads, kStaged ? 1 : 2) hit_kernel(const HitArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint4* tab = reinterpret_cast<uint4*>(smem_raw);
	const uint32_t tid = threadIdx.x, g = tid / kGroupThreads, gtid = tid % kGroupThreads;
	constexpr uint32_t kQCap = kStaged ? kStagedQueueCap : kQueueCap;
	const size_t group_bytes = group_smem_fixed(kQCap) + (kStaged ? (size_t)kTileRecs * a.stride * 4 : 0);
	const GroupSmem sm = group_smem(smem_raw + kTabBytes + g * group_bytes, kQCap);
	for (uint32_t i = tid; i < 8 * 256; i += kGroups * kGroupThreads)
		tab[i] = a.d_tab[i];
	// open / spare blocks of this group: carried over from the previous launch unless a flush or a reset came between
	const uint32_t group_id = blockIdx.x * kGroups + g;
	const uint32_t nb = a.pool.nbins;
	uint32_t* gs = a.pool.gstate + ((size_t)a.ki * a.pool.max_groups + group_id) * gstate_row(nb) + 1; // gs[-1] = epoch, gs[0] = flushes + 1
	const uint32_t gen = a.pool.ctl[CTL_FLUSHES] + 1u; // the full counters: no wrap can make a stale state look current
	const bool active = group_id < a.n_units && group_id < a.pool.max_groups;
	const bool resume = active && gs[-1] == a.pool.epoch && gs[0] == gen;
	Spare sp;
	sp.have = false;
	sp.blk = sp.at = 0;
	if (gtid < kMaxBins) {
		sm.cur[gtid] = kVoid;
		sm.fill[gtid] = kBlkEntries;
		if (resume && gtid < nb) {
			sm.cur[gtid] = gs[1 + gtid];
			sm.fill[gtid] = gs[1 + nb + gtid];
			sm.curpos[gtid] = gs[1 + 2 * nb + gtid];
			sp.blk = gs[1 + 3 * nb + gtid];
			sp.at = gs[1 + 4 * nb + gtid];
			sp.have = true;
		}
	}
	if (gtid == 0) {
		sm.scal[2] = 0;
		if (kStaged) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(sm.mbar)), "r"(1u) : "memory");
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	}
	__syncthreads();
	if (gtid < nb && active && !resume)
		top_up(sm, a, gtid, sp); // open a first block for every bin
	constexpr int kW = 8; // mask words per thread in flight
	uint32_t phase = 0;
	for (uint32_t unit = blockIdx.x * kGroups + g; unit < a.n_units; unit += gridDim.x * kGroups) {
		Unit U;
		U.multi = a.tiles_per_unit > 1;
		if (!U.multi) {
			U.tile0 = unit / a.units_per_tile;
			U.r0 = (unit - U.tile0 * a.units_per_tile) * a.rows_per_unit;
			const uint32_t info = __ldg(a.tile_info + U.tile0);          // rows past the tile's k-mer positions are all zero
			const uint32_t npos = info == kTileFlag ? 0u : min(info, a.npos_max);
			U.nrows = npos > U.r0 ? min(a.rows_per_unit, npos - U.r0) : 0u;
		} else {
			U.tile0 = unit * a.tiles_per_unit;
			U.r0 = 0;
			U.nrows = min(a.tiles_per_unit, a.n_tiles - U.tile0) * a.npos_max;
		}
		const uint32_t nw = U.nrows * 32u;
		const uint32_t* m = a.masks + ((size_t)U.tile0 * a.npos_max + U.r0) * 32u;
		if (a.mask_prefetch && !U.multi) {
			// the next unit's mask rows (128 bytes each) on their way to L2 while this one is processed: the enumeration below
			// otherwise starts every unit with a full DRAM latency that nothing covers (the whole group waits for it)
			const uint32_t nu = unit + gridDim.x * kGroups;
			if (nu < a.n_units) {
				const uint32_t t1 = nu / a.units_per_tile, r1 = (nu - t1 * a.units_per_tile) * a.rows_per_unit;
				const uint32_t rows = min(a.rows_per_unit, a.npos_max - min(r1, a.npos_max));
				const uint32_t* m1 = a.masks + ((size_t)t1 * a.npos_max + r1) * 32u;
				for (uint32_t r = gtid; r < rows; r += kGroupThreads)
					asm volatile("prefetch.global.L2 [%0];" ::"l"(m1 + (size_t)r * 32u));
			}
		}
		if (gtid == 0) {
			sm.scal[0] = 0;
			sm.scal[1] = 0;
			if (kStaged) {
				// everybody finished reading the previous tile (group_sync at the end of the last round): fetch this one
				const uint32_t nrec = min(kTileRecs, a.n_rec - U.tile0 * kTileRecs);
				const uint32_t bytes = nrec * a.stride * 4u;
				const uint32_t* src = a.words + (size_t)U.tile0 * kTileRecs * a.stride;
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(sm.mbar)), "r"(bytes) : "memory");
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sm.tile)),
				             "l"(src), "r"(bytes), "r"(smem_addr(sm.mbar))
				             : "memory");
			}
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
				uint32_t at = atomicAdd(&sm.scal[0], cnt);
				if (at + cnt <= kQCap) {
#pragma unroll
					for (int u = 0; u < kW; u++)
						emit_bits(sm, x[u], w0 + u * kGroupThreads + gtid, at);
				} else {
					sm.scal[1] = 1;
				}
			}
		}
		group_sync(g);
		if (kStaged) { // the tile has landed?
			uint32_t done = 0;
			while (!done)
				asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				             : "=r"(done)
				             : "r"(smem_addr(sm.mbar)), "r"(phase)
				             : "memory");
			phase ^= 1u;
		}
		if (!sm.scal[1]) {
			const uint32_t n = sm.scal[0];
			if (n)
				process_queue<kStaged>(sm, a, tab, n, U, g, gtid, sp);
		} else {
			// skewed data: more candidates than the queue holds -> two mask rows (<= 2048 candidates) per round
			for (uint32_t w0 = 0; w0 < nw; w0 += 64) {
				group_sync(g);
				if (gtid == 0)
					sm.scal[0] = 0;
				group_sync(g);
				const uint32_t w = w0 + gtid;
				const uint32_t x = (gtid < 64 && w < nw) ? __ldg(m + w) : 0u;
				if (x) {
					uint32_t at = atomicAdd(&sm.scal[0], (uint32_t)__popc(x));
					emit_bits(sm, x, w, at);
				}
				group_sync(g);
				const uint32_t n = sm.scal[0];
				if (n)
					process_queue<kStaged>(sm, a, tab, n, U, g, gtid, sp);
			}
		}
		group_sync(g); // the staged tile and the queue are free again
	}
	// keep the open and spare blocks for the next launch; the spare is registered as an empty block
	if (active) {
		group_sync(g);
		if (gtid < nb) {
			if (sp.have) {
				uint32_t* list = a.pool.slice_blocks + (size_t)(a.ki * nb + gtid) * a.pool.slice_cap;
				if (sp.at < a.pool.slice_cap)
					list[sp.at] = (sp.blk < a.pool.n_blocks) ? (sp.blk << 9) : kVoid;
			}
			gs[1 + gtid] = sm.cur[gtid];
			gs[1 + nb + gtid] = sm.fill[gtid];
			gs[1 + 2 * nb + gtid] = sm.curpos[gtid];
			gs[1 + 3 * nb + gtid] = sp.blk;
			gs[1 + 4 * nb + gtid] = sp.at;
		}
		if (gtid == 0) {
			gs[-1] = a.pool.epoch;
			gs[0] = sp.have ? gen : 0u;
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
				process_piece_k(p + 1, min(__ldg(p), (stride - 1u) * 16u), 0, k, T, ctr_k, rBits, sBits); // F1 was counted by the scan kernel
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


// order / n_order: the slices to flush, in pipeline order (nullptr = all n_slices); the log of every other slice is dropped
__global__ void __launch_bounds__(kApplyThreads) apply_kernel(const Pool P, uint32_t* __restrict__ counters, int force, uint32_t reserve_blocks,
    const uint32_t* __restrict__ order, uint32_t n_order)
{
	const uint32_t kApplyAhead = P.ahead;
	const uint32_t NS = order ? n_order : P.n_slices;
	// ---- decide (every CTA reads the same values: they were written by earlier kernels) ----
	const uint32_t next = P.ctl[CTL_NEXT];
	const uint32_t used = min(next, P.n_blocks);
	const uint32_t state = P.ctl[CTL_STATE];
	if (force == 2) {
		// after the fused kernel's first pass: tiles were deferred (pool exhausted before the sketch was materialised) or flagged
		// for the fallback kernel, or a materialised sketch's pool is 3/4 full (full = direct increments into HBM, ~8x slower)
		const bool flush = P.ctl[CTL_NDEFER] != 0 || P.ctl[CTL_NFLAG] != 0 || (state == 1 && used > P.n_blocks / 4 * 3);
		if (!flush)
			return;
	} else if (!force) {
		const unsigned long long need = *P.cand / (kBlkEntries - kBlkSlack) + reserve_blocks; // closed blocks hold > 256 - slack entries
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
		for (uint32_t si = 0; si < NS; si++) {
			const uint32_t s = order ? order[si] : si;
			if (si >= kApplyAhead) { // L2 footprint: do not run more than kApplyAhead slices ahead of the appliers
				if (rtid == 0)
					while (ld_acquire(P.apply_done + si - kApplyAhead) < gridDim.x)
						__nanosleep(20);
				role_sync(role);
			}
			uint4* p = reinterpret_cast<uint4*>(counters + (size_t)(s / P.nbins) * per_k + (size_t)(s % P.nbins) * slice_len);
			if (zero_mode) {
				for (size_t i = a0 + rtid; i < a1; i += nrt)
					p[i] = make_uint4(0u, 0u, 0u, 0u);
				role_sync(role);
				if (rtid == 0) {
					__threadfence();
					atomicAdd(P.zero_done + si, 1u);
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
		// One warp per log block.  The first block of the NEXT slice is loaded (list entry -> 256 entries) before
		// the warp waits for the current slice to be staged, so the loads' latency is off the per-slice chain.
		const uint32_t lane = rtid & 31u, warp = rtid >> 5, nwarp = nrt / 32;
		constexpr int kV = kBlkEntries / 32;
		const uint32_t j0 = blockIdx.x * nwarp + warp, jstep = gridDim.x * nwarp;
		auto load_block = [&](uint32_t s, uint32_t j, uint32_t nb, uint32_t (&v)[kV]) {
			uint32_t le = kVoid;
			if (j < nb)
				le = __ldcs(P.slice_blocks + (size_t)s * P.slice_cap + j);
			const uint32_t f = le == kVoid ? 0u : min(le & 511u, kBlkEntries);
			const uint32_t* e = P.entries + (size_t)(le >> 9) * kBlkEntries;
#pragma unroll
			for (int u = 0; u < kV; u++)
				v[u] = lane + 32u * u < f ? __ldcs(e + lane + 32u * u) : kVoid;
		};
		uint32_t v[kV], vn[kV];
		auto slice_at = [&](uint32_t si) { return order ? order[si] : si; };
		uint32_t nb = NS ? min(P.slice_nblk[slice_at(0)], P.slice_cap) : 0u;
		if (NS)
			load_block(slice_at(0), j0, nb, v);
		for (uint32_t si = 0; si < NS; si++) {
			const uint32_t s = slice_at(si);
			const uint32_t sn = si + 1 < NS ? slice_at(si + 1) : 0u;
			const uint32_t nbn = si + 1 < NS ? min(P.slice_nblk[sn], P.slice_cap) : 0u;
			if (si + 1 < NS)
				load_block(sn, j0, nbn, vn);
			if (nb) {
				if (zero_mode) {
					if (rtid == 0)
						while (ld_acquire(P.zero_done + si) < gridDim.x)
							__nanosleep(20);
					role_sync(role);
				}
				uint32_t* ctr = counters + (size_t)(s / P.nbins) * per_k;
#pragma unroll
				for (int u = 0; u < kV; u++)
					if (v[u] != kVoid)
						atomicAdd(ctr + v[u], 1u); // RED.ADD, L2 resident
				for (uint32_t j = j0 + jstep; j < nb; j += jstep) { // more blocks than warps: the rest, unpipelined
					load_block(s, j, nb, v);
#pragma unroll
					for (int u = 0; u < kV; u++)
						if (v[u] != kVoid)
							atomicAdd(ctr + v[u], 1u);
				}
			}
			role_sync(role);
			if (rtid == 0)
				atomicAdd(P.apply_done + si, 1u);
#pragma unroll
			for (int u = 0; u < kV; u++)
				v[u] = vn[u];
			nb = nbn;
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


// ------------------------------------------------------------------------------------------------
// multi-GPU reduction over peer memory (NVLink): apply_owned_kernel
// ------------------------------------------------------------------------------------------------
// Rank r owns the sketch slices s with s % world == r.  Every rank's hit log (block lists + pool) is mapped into every
// other rank's address space (CUDA IPC, ntc_peer_attach); this kernel zeroes the OWNED slices of the local sketch and
// RED.ADDs into them the log entries of those slices from ALL ranks' logs -- the remote ones are read straight through
// NVLink with coalesced 1 KB block loads, two blocks per warp in flight.  The fused compute + collective step of the
// reduction: no staging buffers, no host planning, no NCCL on the data path (NCCL only carries the small all-reduces
// around it, which also are the barriers: the status all-reduce before -- every log is complete -- and the histogram
// all-reduce after -- every rank has finished reading its peers' pools).
// status[status_idx] != 0 (device memory, all-reduced): some rank's log is incomplete -- do nothing (the host falls
// back to the dense reduction once it has read the flag).
__global__ void __launch_bounds__(kApplyThreads) apply_owned_kernel(const Pool P, const PeerLogs G, uint32_t* __restrict__ counters,
    const uint32_t* __restrict__ order, uint32_t n_order, const long long* __restrict__ status, uint32_t status_idx)
{
	if (status && status[status_idx] != 0)
		return;
	const uint32_t kApplyAhead = P.ahead;
	const uint32_t role = threadIdx.x / (kApplyThreads / 2), rtid = threadIdx.x % (kApplyThreads / 2), nrt = kApplyThreads / 2;
	const size_t slice_len = (size_t)1 << P.bin_shift;
	const size_t per_k = (size_t)2 << P.rBits;
	if (role == 0) {
		const size_t n16 = slice_len / 4;
		const size_t a0 = n16 * blockIdx.x / gridDim.x, a1 = n16 * (blockIdx.x + 1) / gridDim.x;
		for (uint32_t si = 0; si < n_order; si++) {
			const uint32_t s = order[si];
			if (si >= kApplyAhead) {
				if (rtid == 0)
					while (ld_acquire(P.apply_done + si - kApplyAhead) < gridDim.x)
						__nanosleep(20);
				role_sync(role);
			}
			uint4* p = reinterpret_cast<uint4*>(counters + (size_t)(s / P.nbins) * per_k + (size_t)(s % P.nbins) * slice_len);
			for (size_t i = a0 + rtid; i < a1; i += nrt)
				p[i] = make_uint4(0u, 0u, 0u, 0u);
			role_sync(role);
			if (rtid == 0) {
				__threadfence();
				atomicAdd(P.zero_done + si, 1u);
			}
		}
	} else {
		const uint32_t lane = rtid & 31u, warp = rtid >> 5, nwarp = nrt / 32;
		constexpr int kV = kBlkEntries / 32;
		const uint32_t j0 = blockIdx.x * nwarp + warp, jstep = gridDim.x * nwarp;
		for (uint32_t si = 0; si < n_order; si++) {
			const uint32_t s = order[si];
			uint32_t* ctr = counters + (size_t)(s / P.nbins) * per_k;
			// The blocks of slice s in ALL logs form one index space (peer q', block j), walked by all warps together: a loop over
			// the peers with a loop over blocks inside costs one NVLink round trip per (slice, peer) and leaves most warps idle
			// (N = 8, 64 slices per k: 0.53 ms against 0.36 ms).  The peers are visited starting behind this rank, so that the
			// ranks do not all pull from rank 0 at the same time.
			uint32_t first[kMaxPeers + 1];
			first[0] = 0;
#pragma unroll
			for (uint32_t i = 0; i < kMaxPeers; i++) {
				const uint32_t q = (G.self + 1u + i) % G.n;
				first[i + 1] = first[i] + (i < G.n ? min(__ldcg(G.slice_nblk[q] + s), P.slice_cap) : 0u);
			}
			const uint32_t total = first[kMaxPeers];
			auto locate = [&](uint32_t g, const uint32_t*& list, const uint32_t*& ent, uint32_t& j) {
				uint32_t i = 0;
#pragma unroll
				for (uint32_t t = 1; t < kMaxPeers; t++)
					i += g >= first[t] ? 1u : 0u;
				const uint32_t q = (G.self + 1u + i) % G.n;
				list = G.slice_blocks[q] + (size_t)s * P.slice_cap;
				ent = G.entries[q];
				j = g - first[i];
			};
			constexpr int kFly = 4; // blocks in flight per warp: a list entry, then 1 KB of entries, each an NVLink round trip
			// The loads of the first round are requested BEFORE the warp waits for the slice to be zeroed -- they do not depend on it,
			// and with one round per warp and slice (9 k blocks, 2.4 k warps) the three remote round trips (block counts, list
			// entries, entries) would otherwise be paid once per slice, one after the other, behind the zero-fill.
			bool waited = false;
			for (uint32_t g = j0; g < total || !waited; g += kFly * jstep) {
				const uint32_t* eb[kFly];
				uint32_t fl[kFly];
				uint32_t le[kFly];
#pragma unroll
				for (int f = 0; f < kFly; f++) {
					const uint32_t gf = g + f * jstep;
					le[f] = kVoid;
					eb[f] = nullptr;
					if (gf < total) {
						const uint32_t* list;
						uint32_t j;
						locate(gf, list, eb[f], j);
						le[f] = __ldcg(list + j);
					}
				}
				uint32_t v[kFly][kV];
#pragma unroll
				for (int f = 0; f < kFly; f++) {
					fl[f] = le[f] == kVoid ? 0u : min(le[f] & 511u, kBlkEntries);
					const uint32_t* e = eb[f] + (size_t)(le[f] >> 9) * kBlkEntries;
#pragma unroll
					for (int u = 0; u < kV; u++)
						v[f][u] = lane + 32u * u < fl[f] ? __ldcg(e + lane + 32u * u) : kVoid;
				}
				if (!waited) { // every applier thread passes here exactly once per slice
					if (rtid == 0)
						while (ld_acquire(P.zero_done + si) < gridDim.x)
							__nanosleep(20);
					role_sync(role);
					waited = true;
				}
#pragma unroll
				for (int f = 0; f < kFly; f++)
#pragma unroll
					for (int u = 0; u < kV; u++)
						if (v[f][u] != kVoid)
							atomicAdd(ctr + v[f][u], 1u);
			}
			role_sync(role);
			if (rtid == 0)
				atomicAdd(P.apply_done + si, 1u);
		}
	}
	// the pool is NOT reset here: peers may still be reading it (ntc_reset does, after the histogram all-reduce)
}

// F1 per k and "my log is not complete" as int64, ready for one all-reduce (sum)
__global__ void log_status_kernel(const Pool P, const unsigned long long* __restrict__ f1, uint32_t nK, int host_ok, long long* __restrict__ out)
{
	if (threadIdx.x < nK)
		out[threadIdx.x] = (long long)f1[threadIdx.x];
	if (threadIdx.x == 0) {
		const bool ok = host_ok && P.ctl[CTL_STATE] == 0 && P.ctl[CTL_DIRECT] == 0 && P.ctl[CTL_NEXT] <= P.n_blocks && P.ctl[CTL_NDEFER] == 0;
		out[nK] = ok ? 0 : 1;
	}
}

} // namespace

constexpr int kStagedGroups = 3, kPlainGroups = 2;

static size_t hit_smem(bool staged, uint32_t stride)
{
	const size_t per_group = group_smem_fixed(staged ? kStagedQueueCap : kQueueCap) + (staged ? (size_t)kTileRecs * stride * 4 : 0);
	return kTabBytes + (staged ? kStagedGroups : kPlainGroups) * per_group;
}

// staged: one unit = one whole tile and three groups with their tiles fit the 227 KB of an SM
bool hit_can_stage(uint32_t stride, uint32_t units_per_tile, uint32_t tiles_per_unit)
{
	return units_per_tile == 1 && tiles_per_unit == 1 && hit_smem(true, stride) <= 232448;
}

unsigned hit_groups_per_sm(bool staged) { return staged ? kStagedGroups : 2 * kPlainGroups; }

cudaError_t launch_hit(const HitArgs& a, bool staged, unsigned ctas)
{
	const size_t smem = hit_smem(staged, a.stride);
	cudaError_t e;
	// the opt-in shared-memory limit is a property of (function, device): set it when it has to grow, not on every launch
	static size_t smem_set[2][64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	const bool known = dev >= 0 && dev < 64;
	const bool raise = !known || smem > smem_set[staged ? 1 : 0][dev];
	if (staged) {
		auto kern = hit_kernel<true, kStagedGroups>;
		if (raise && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
			return e;
		kern<<<ctas, kStagedGroups * kGroupThreads, smem, a.stream>>>(a);
	} else {
		auto kern = hit_kernel<false, kPlainGroups>;
		if (raise && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
			return e;
		kern<<<ctas, kPlainGroups * kGroupThreads, smem, a.stream>>>(a);
	}
	if (raise && known)
		smem_set[staged ? 1 : 0][dev] = smem;
	return cudaGetLastError();
}

// Start of a batch: the candidate count and the flagged-tile count back to zero -- as one kernel instead of two cudaMemsetAsync.
// Experiment only (NTC_CLEAR_MEMSET=0): it was meant to keep the tiny memsets out of the copy engine's queue, but measured far slower on the
// host side (ntc_api.cu).
__global__ void batch_clear_kernel(unsigned long long* cand, uint32_t* ctl)
{
	if (threadIdx.x == 0) {
		*cand = 0ull;
		ctl[CTL_NFLAG] = 0u;
	}
}

cudaError_t launch_batch_clear(const Pool& pool, cudaStream_t st)
{
	batch_clear_kernel<<<1, 32, 0, st>>>(pool.cand, pool.ctl);
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
	// one grid size for both cooperative kernels (every CTA must be resident: their roles wait for each other)
	int occ = 0, occ2 = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, apply_kernel, (int)kApplyThreads, 0) != cudaSuccess || occ < 1)
		occ = 1;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, apply_owned_kernel, (int)kApplyThreads, 0) != cudaSuccess || occ2 < 1)
		occ2 = 1;
	occ = occ < occ2 ? occ : occ2;
	return n_sm * (occ > 2 ? 2 : occ);
}

cudaError_t launch_apply(const Pool& pool, uint32_t* counters, int force, uint32_t reserve_blocks, unsigned grid, cudaStream_t st,
    const uint32_t* d_order, uint32_t n_order)
{
	// Cooperative launch: the stagers and appliers of different CTAs wait for each other through counters in global
	// memory, so every CTA of the grid must be resident at once -- also when another context's kernels share the GPU.
	Pool p = pool;
	static const bool plain = getenv("NTC_APPLY_COOP") && atoi(getenv("NTC_APPLY_COOP")) == 0; // experiment: ordinary launch
	if (plain) {
		apply_kernel<<<grid, kApplyThreads, 0, st>>>(p, counters, force, reserve_blocks, d_order, n_order);
		return cudaGetLastError();
	}
	void* args[] = { &p, &counters, &force, &reserve_blocks, &d_order, &n_order };
	return cudaLaunchCooperativeKernel((const void*)apply_kernel, dim3(grid), dim3(kApplyThreads), args, 0, st);
}

// ------------------------------------------------------------------------------------------------
// hit-log exchange (multi-GPU): export the blocks of some slices into contiguous buffers; import blocks
// received from other ranks into the free part of the pool
// ------------------------------------------------------------------------------------------------
// runs: [n_runs][2] = {slice, first output / input block of the run}; run r covers blocks [runs[r][1], runs[r+1][1])
__device__ __forceinline__ uint32_t find_run(const uint32_t* __restrict__ runs, uint32_t n_runs, uint32_t blk)
{
	uint32_t lo = 0, hi = n_runs; // last run whose first block <= blk
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) / 2;
		if (runs[2 * mid + 1] <= blk)
			lo = mid;
		else
			hi = mid;
	}
	return lo;
}

namespace {
__global__ void __launch_bounds__(256) export_kernel(const Pool P, const uint32_t* __restrict__ runs, uint32_t n_runs, uint32_t n_out,
    uint32_t* __restrict__ out)
{
	const uint32_t lane = threadIdx.x & 31u;
	for (uint32_t o = blockIdx.x * 8 + (threadIdx.x >> 5); o < n_out; o += gridDim.x * 8) {
		const uint32_t r = find_run(runs, n_runs, o);
		const uint32_t s = runs[2 * r], j = o - runs[2 * r + 1];
		const uint32_t le = j < P.slice_cap ? P.slice_blocks[(size_t)s * P.slice_cap + j] : kVoid;
		const uint32_t f = le == kVoid ? 0u : min(le & 511u, kBlkEntries);
		if (lane == 0)
			out[(size_t)o * kWireBlkWords + kBlkEntries] = f; // trailer: valid entries
		const uint32_t* e = P.entries + (size_t)(le >> 9) * kBlkEntries;
		for (uint32_t i = lane; i < f; i += 32)
			out[(size_t)o * kWireBlkWords + i] = e[i];
	}
}

__global__ void __launch_bounds__(256) import_kernel(const Pool P, const uint32_t* __restrict__ runs, uint32_t n_runs, uint32_t n_in,
    uint32_t base_blk, const uint32_t* __restrict__ in)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0)
		P.ctl[CTL_NEXT] = base_blk + n_in; // the imported blocks now belong to the pool
	if (i >= n_in)
		return;
	const uint32_t s = runs[2 * find_run(runs, n_runs, i)];
	const uint32_t f = min(in[(size_t)i * kWireBlkWords + kBlkEntries], kBlkEntries);
	if (f == 0)
		return;
	const uint32_t at = atomicAdd(P.slice_nblk + s, 1u);
	if (at < P.slice_cap)
		P.slice_blocks[(size_t)s * P.slice_cap + at] = ((base_blk + i) << 9) | f;
	else
		P.ctl[CTL_DIRECT] = 2u; // cannot happen: the host checks the capacities before importing
}
} // namespace

cudaError_t launch_export(const Pool& pool, const uint32_t* d_runs, uint32_t n_runs, uint32_t n_out, uint32_t* d_out, cudaStream_t st)
{
	if (n_out == 0)
		return cudaSuccess;
	const unsigned grid = (n_out + 7) / 8 < 148 * 8 ? (n_out + 7) / 8 : 148 * 8;
	export_kernel<<<grid, 256, 0, st>>>(pool, d_runs, n_runs, n_out, d_out);
	return cudaGetLastError();
}

cudaError_t launch_import(const Pool& pool, const uint32_t* d_runs, uint32_t n_runs, uint32_t n_in, uint32_t base_blk, const uint32_t* d_in,
    cudaStream_t st)
{
	import_kernel<<<(n_in + 255) / 256 + (n_in == 0), 256, 0, st>>>(pool, d_runs, n_runs, n_in, base_blk, d_in);
	return cudaGetLastError();
}

// counter-value histogram (v >= 1) of a list of slices in ONE launch: CTA = 16 K consecutive counters
namespace {
constexpr uint32_t kHistChunk = 16384, kHistBins = 2048;
__global__ void __launch_bounds__(256) hist_slices_kernel(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ order, uint32_t bin_shift,
    uint32_t nbins, uint32_t rBits, uint32_t* __restrict__ p_hist)
{
	__shared__ uint32_t sh[kHistBins];
	for (uint32_t i = threadIdx.x; i < kHistBins; i += blockDim.x)
		sh[i] = 0;
	__syncthreads();
	const uint64_t slice_len = (uint64_t)1 << bin_shift, per_k = (uint64_t)2 << rBits;
	const uint32_t chunk = (uint32_t)min((uint64_t)kHistChunk, slice_len);
	const uint32_t chunks_per_slice = (uint32_t)(slice_len / chunk);
	const uint32_t s = order ? order[blockIdx.x / chunks_per_slice] : blockIdx.x / chunks_per_slice; // no list: every slice
	const uint64_t first = (uint64_t)(s / nbins) * per_k + (uint64_t)(s % nbins) * slice_len + (uint64_t)(blockIdx.x % chunks_per_slice) * chunk;
	uint32_t* ph = p_hist + (first >> rBits) * 65536; // a chunk lies inside one table (chunk <= 2^rBits: the host checks)
	auto count = [&](uint32_t v) {
		v &= 0xFFFFu; // uint16_t wrap of ntcard.cpp:143
		if (v) {
			if (v < kHistBins)
				atomicAdd(&sh[v], 1u);
			else
				atomicAdd(&ph[v], 1u);
		}
	};
	if (chunk % 4 == 0) {
		const uint4* src = reinterpret_cast<const uint4*>(counters + first);
		for (uint32_t i = threadIdx.x; i < chunk / 4; i += blockDim.x) {
			const uint4 v = __ldcs(src + i);
			if (v.x | v.y | v.z | v.w) {
				count(v.x);
				count(v.y);
				count(v.z);
				count(v.w);
			}
		}
	} else {
		for (uint32_t i = threadIdx.x; i < chunk; i += blockDim.x)
			count(counters[first + i]);
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < kHistBins; i += blockDim.x)
		if (sh[i])
			atomicAdd(&ph[i], sh[i]);
}
} // namespace

cudaError_t launch_hist_slices(const Pool& pool, const uint32_t* counters, const uint32_t* d_order, uint32_t n_order, uint32_t* d_phist, cudaStream_t st)
{
	if (!d_order)
		n_order = pool.n_slices; // all of the sketch
	if (n_order == 0)
		return cudaSuccess;
	const uint64_t slice_len = (uint64_t)1 << pool.bin_shift;
	const uint64_t chunk = slice_len < kHistChunk ? slice_len : kHistChunk;
	if (chunk > ((uint64_t)1 << pool.rBits))
		return cudaErrorInvalidValue;
	hist_slices_kernel<<<(unsigned)(n_order * (slice_len / chunk)), 256, 0, st>>>(counters, d_order, pool.bin_shift, pool.nbins, pool.rBits, d_phist);
	return cudaGetLastError();
}

cudaError_t launch_apply_owned(const Pool& pool, const PeerLogs& logs, uint32_t* counters, const uint32_t* d_order, uint32_t n_order,
    const long long* d_status, uint32_t status_idx, unsigned grid, cudaStream_t st)
{
	Pool p = pool;
	PeerLogs g = logs;
	void* args[] = { &p, &g, &counters, &d_order, &n_order, &d_status, &status_idx };
	return cudaLaunchCooperativeKernel((const void*)apply_owned_kernel, dim3(grid), dim3(kApplyThreads), args, 0, st);
}

cudaError_t launch_log_status(const Pool& pool, const unsigned long long* d_f1, uint32_t nK, int host_ok, long long* d_out, cudaStream_t st)
{
	log_status_kernel<<<1, 32, 0, st>>>(pool, d_f1, nK, host_ok, d_out);
	return cudaGetLastError();
}

} // namespace pl
} // namespace ntc
