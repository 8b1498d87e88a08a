// fused_kernel.cuh -- the sketch kernel: bit-sliced sampling filter + full hash + hit-log append in ONE launch.
//
// What it replaces: scan_kernel -> masks in HBM -> hit_kernel (pipeline.h).  There the packed reads crossed DRAM twice, 190 MB
// of mask words per 10 M reads were written and read back, the hit kernel spent ~40 % of its instructions finding the 1-in-64
// set bits again, and the two kernels -- one bound by the integer ALU pipe, one by latency -- ran one after the other.
// Here a warp keeps its tile (1024 records) to itself from the first load to the last log entry:
//
//   scan     as in scan_kernel.cuh (ntHashIterator / NTC64 roll, ntHashIterator.hpp:59-86, nthash.hpp:242-279, reduced to the
//            upper 31-bit rings, bit-sliced over 32 records per lane); the 32-slot mask of a position never leaves the
//            registers: its set bits go to a lane-private queue in shared memory as 16-bit words (k-mer start << 5 | slot)
//   hash     at the end of the tile every lane re-hashes its own candidates in full 64 bits from the packed bases (L2 hits:
//            the warp read them a few microseconds ago), applies ntComp (ntcard.cpp:132-145) -> counter index or nothing;
//            the indices are parked in the (now dead) plane ring
//   append   the warp owns one open log block per sketch slice.  Lane b counts the round's hits of slice b, the warp takes all
//            the new blocks it needs with ONE atomicAdd on the pool cursor, then the indices are stored.  All or nothing:
//            if the pool cannot serve the round, nothing has been appended yet, and the warp either increments the sketch
//            directly (legal once it is materialised: CTL_STATE = 1) or DEFERS the tile from its current position
//            (tile_info = kTileDefer | p).  The host always enqueues  pass 0 -> conditional flush -> pass 1 (deferred tiles
//            only; leaves at once when there are none), so exactness never depends on a capacity.
//
// While one warp of an SM sub-partition is in its hash/append phase (latency-bound) the other one scans (ALU-bound).
// A lane's queue holds qlane candidates; when a tile has more (skewed data: low-complexity reads whose k-mer is sampled), the
// round ends at the first position that did not fit and the tile is scanned again from there.
//
// No tensor cores: there is no dense contraction; the kernel is bound by the 16-lane integer ALU pipe.
#pragma once
#include <cuda_runtime.h>

#include "bitslice_core.cuh"
#include "hit_hash.cuh"
#include "pipeline.h"
#include "scan_kernel.cuh"
#include "sketch_common.cuh"

namespace ntc {
namespace pl {

constexpr uint32_t kWarpStateArrays = 7; // cur, fill, curpos, wcnt, wbase, wat, wcursor: nbins words each
constexpr int kFusedBatch = 8;                     // candidates per lane whose loads are in flight together

__host__ __device__ constexpr size_t fused_queue_bytes(uint32_t qlane) { return (((size_t)qlane * 64) + 15) & ~(size_t)15; }
__host__ __device__ constexpr size_t fused_warp_bytes(uint32_t ring, uint32_t qlane, uint32_t nbins)
{
	return (size_t)(ring + 3u) * 256u + fused_queue_bytes(qlane) + (size_t)kWarpStateArrays * nbins * 4u;
}

struct WarpSmem {
	uint2* planes;      // [ring + 3][32]
	uint32_t* hq;       // aliases planes: counter indices of the round's candidates, [qlane][32]
	uint16_t* queue;    // [qlane][32]
	uint32_t *cur, *fill, *curpos; // the warp's open block of every slice of this k (persist over tiles and launches)
	uint32_t *wcnt, *wbase, *wat, *wcursor; // per round
};

__device__ __forceinline__ WarpSmem warp_smem(unsigned char* base, uint32_t ring, uint32_t qlane, uint32_t nbins)
{
	WarpSmem w;
	w.planes = reinterpret_cast<uint2*>(base);
	w.hq = reinterpret_cast<uint32_t*>(base);
	w.queue = reinterpret_cast<uint16_t*>(base + (size_t)(ring + 3u) * 256u);
	w.cur = reinterpret_cast<uint32_t*>(base + (size_t)(ring + 3u) * 256u + fused_queue_bytes(qlane));
	w.fill = w.cur + nbins;
	w.curpos = w.fill + nbins;
	w.wcnt = w.curpos + nbins;
	w.wbase = w.wcnt + nbins;
	w.wat = w.wbase + nbins;
	w.wcursor = w.wat + nbins;
	return w;
}

template <int KM, int S, int U> struct FusedBlock {
	static __device__ __forceinline__ void run(bs::State& st, const uint2* __restrict__ pin, const uint2* __restrict__ pout, uint32_t (&m)[kScanBlock])
	{
		const uint2 in = pin[U * 32];
		const uint2 out = pout[U * 32];
		bs::step<KM, U>(st, in.x, in.y, out.x, out.y);
		m[U] = bs::sampled_mask<U, S>(st);
		FusedBlock<KM, S, U + 1>::run(st, pin, pout, m);
	}
};
template <int KM, int S> struct FusedBlock<KM, S, kScanBlock> {
	static __device__ __forceinline__ void run(bs::State&, const uint2* __restrict__, const uint2* __restrict__, uint32_t (&)[kScanBlock]) {}
};

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
		if (lane >= (uint32_t)d)
			v += t;
	}
	return v;
}

// The hash + append phase of one round of one tile.  Returns false when the pool could not serve the round while the sketch is
// not materialised (nothing was appended: the caller defers the tile).
// Not inlined, so that its registers do not add to the scan loop's; the shared-memory pointers are rebuilt from the kernel's
// dynamic shared array here, which keeps them shared-space accesses (LDS / STS) instead of generic ones.
// This phase is bound by the load/store unit (L1 wavefronts of the gathers and of the scattered log stores, the table
// lookups), so it is written to need few of them: 128-bit gathers, and the slot of a hit inside its slice's block is
// assigned with a warp match instead of shared-memory atomics.
static __device__ __noinline__ bool fused_drain(const FusedArgs& a, uint32_t tile, uint32_t cnt, uint32_t p_end, bool mixed, uint32_t lane, uint32_t warp)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const Pool& P = a.pool;
	const uint32_t nb = P.nbins, k = a.L.k;
	const uint4* tab = reinterpret_cast<const uint4*>(smem_raw);
	const WarpSmem sm = warp_smem(smem_raw + kTabBytes + (size_t)warp * fused_warp_bytes(a.L.ring, a.qlane, nb), a.L.ring, a.qlane, nb);
	uint32_t maxcnt = cnt;
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		maxcnt = max(maxcnt, __shfl_xor_sync(0xFFFFFFFFu, maxcnt, d));
	if (maxcnt == 0)
		return true;
	// cursor of every slice: the fill of the warp's open block (a full block when there is none); wcnt keeps the start
	for (uint32_t b = lane; b < nb; b += 32) {
		const uint32_t f = sm.cur[b] != kVoid ? sm.fill[b] : kBlkEntries;
		sm.wcursor[b] = f;
		sm.wcnt[b] = f;
	}
	__syncwarp();
	const HashK K = a.hk;
	const uint32_t rBits = P.rBits, S = a.sBits, bin_shift = P.bin_shift, n_rec = a.n_rec, ngroups = a.stride >> 2;
	const uint4* __restrict__ recs = reinterpret_cast<const uint4*>(a.words);
	const uint32_t lt = (1u << lane) - 1u;
	// ---- candidates -> counter index + slot in the slice's block (parked in hq / queue) ---------------------------
	for (uint32_t i0 = 0; i0 < maxcnt; i0 += kFusedBatch) {
		HitLoadV h[kFusedBatch];
		uint32_t pp[kFusedBatch];
		const uint4* rp[kFusedBatch];
		bool val[kFusedBatch];
#pragma unroll
		for (int u = 0; u < kFusedBatch; u++) {
			const uint32_t i = i0 + u;
			const uint32_t e = i < cnt ? sm.queue[i * 32u + lane] : 0u;
			const uint32_t p = e >> 5, rec = tile * kTileRecs + (e & 31u) * 32u + lane;
			pp[u] = p;
			val[u] = i < cnt && p < p_end && rec < n_rec; // past the round's end / slots past the end of the batch
			rp[u] = recs + (uint64_t)min(rec, n_rec - 1u) * ngroups;
			if (val[u])
				h[u] = hit_issue_v(rp[u], ngroups, 1u + (p >> 4));
		}
#pragma unroll
		for (int u = 0; u < kFusedBatch; u++) {
			const uint32_t i = i0 + u;
			uint32_t idx = kVoid;
			// positions past the end of a record of a mixed-length tile ran on its padding (scan_kernel.cuh)
			if (val[u] && (!mixed || pp[u] + k <= __ldg(reinterpret_cast<const uint32_t*>(rp[u]))))
				idx = hash_kmer_v(K, tab, rBits, S, h[u], rp[u], ngroups, pp[u]);
			const uint32_t b = idx == kVoid ? 0xFFFFu : idx >> bin_shift;
			const uint32_t peers = __match_any_sync(0xFFFFFFFFu, b);
			const uint32_t rank = (uint32_t)__popc(peers & lt);
			uint32_t slot = 0;
			if (idx != kVoid)
				slot = sm.wcursor[b] + rank;
			__syncwarp();
			if (idx != kVoid && rank == 0)
				sm.wcursor[b] = slot + (uint32_t)__popc(peers);
			__syncwarp();
			if (i < maxcnt) {
				sm.hq[i * 32u + lane] = idx;
				sm.queue[i * 32u + lane] = (uint16_t)slot;
			}
		}
	}
	__syncwarp();
	// ---- lane b: blocks for slice b ------------------------------------------------------------------------------
	bool ok = true;
	uint32_t need_[2], f_[2], nb_[2], base_[2];
#pragma unroll
	for (int j = 0; j < 2; j++) {
		const uint32_t b = lane + 32u * j;
		need_[j] = f_[j] = nb_[j] = base_[j] = 0;
		if (32u * j >= nb) // warp-uniform
			continue;
		const bool mine = b < nb;
		const uint32_t f = mine ? sm.wcnt[b] : kBlkEntries;
		const uint32_t need = mine ? sm.wcursor[b] - f : 0u;
		const uint32_t room = kBlkEntries - f;
		const uint32_t nbj = need > room ? (need - room + kBlkEntries - 1) / kBlkEntries : 0u;
		const uint32_t incl = warp_incl_scan(nbj, lane);
		const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
		need_[j] = need;
		f_[j] = f;
		nb_[j] = nbj;
		if (total) {
			uint32_t base = 0;
			if (lane == 0)
				base = atomicAdd(P.ctl + CTL_NEXT, total);
			base = __shfl_sync(0xFFFFFFFFu, base, 0);
			if ((uint64_t)base + total > P.n_blocks)
				ok = false;
			base_[j] = base + incl - nbj;
		}
	}
	if (!ok) {
		// The pool is exhausted (the blocks just taken past its end are simply lost).  Materialised sketch: increment it
		// directly.  Otherwise: nothing of this round has been appended -- defer.
		if (ld_volatile(P.ctl + CTL_STATE) != 1u)
			return false;
		for (uint32_t i = 0; i < maxcnt; i++) {
			const uint32_t idx = sm.hq[i * 32u + lane];
			if (idx != kVoid)
				atomicAdd(a.ctr_k + idx, 1u);
		}
		if (lane == 0)
			P.ctl[CTL_DIRECT] = 1u; // statistics; also marks the log as incomplete for the exchange
		__syncwarp();
		return true;
	}
#pragma unroll
	for (int j = 0; j < 2; j++) {
		const uint32_t b = lane + 32u * j;
		if (b < nb) {
			uint32_t at = 0;
			if (nb_[j])
				at = atomicAdd(P.slice_nblk + (a.ki * nb + b), nb_[j]); // slice_cap == n_blocks: the list cannot overflow before the pool
			sm.wbase[b] = base_[j];
			sm.wat[b] = at;
		}
	}
	__syncwarp();
	// ---- append ----------------------------------------------------------------------------------------------------
	for (uint32_t i = 0; i < maxcnt; i++) {
		const uint32_t idx = sm.hq[i * 32u + lane];
		if (idx != kVoid) {
			const uint32_t b = idx >> bin_shift;
			const uint32_t slot = sm.queue[i * 32u + lane];
			uint32_t blk, off;
			if (slot < kBlkEntries) {
				blk = sm.cur[b];
				off = slot;
			} else {
				blk = sm.wbase[b] + ((slot - kBlkEntries) >> 8);
				off = (slot - kBlkEntries) & (kBlkEntries - 1u);
			}
			P.entries[(size_t)blk * kBlkEntries + off] = idx;
		}
	}
	__syncwarp();
	// ---- lane b: the block lists of slice b --------------------------------------------------------------------------
#pragma unroll
	for (int j = 0; j < 2; j++) {
		const uint32_t b = lane + 32u * j;
		if (b < nb && need_[j]) {
			uint32_t* list = P.slice_blocks + (size_t)(a.ki * nb + b) * P.slice_cap;
			const uint32_t total = f_[j] + need_[j], cu = sm.cur[b];
			if (cu != kVoid)
				list[sm.curpos[b]] = (cu << 9) | min(total, kBlkEntries);
			if (nb_[j]) {
				const uint32_t at = sm.wat[b];
				for (uint32_t t = 0; t < nb_[j]; t++)
					list[at + t] = ((base_[j] + t) << 9) | min(kBlkEntries, total - kBlkEntries * (t + 1u));
				sm.cur[b] = base_[j] + nb_[j] - 1u;
				sm.curpos[b] = at + nb_[j] - 1u;
				sm.fill[b] = total - kBlkEntries * nb_[j];
			} else {
				sm.fill[b] = total;
			}
		}
	}
	__syncwarp();
	return true;
}

template <int KM, int S>
__global__ void __launch_bounds__(kScanThreads, 1) fused_kernel(const __grid_constant__ FusedArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const ScanLaunch& L = a.L;
	const Pool& P = a.pool;
	if (a.pass == 1 && P.ctl[CTL_NDEFER] == 0) // nothing was deferred: the second pass has nothing to do
		return;
	uint4* tab = reinterpret_cast<uint4*>(smem_raw);
	for (uint32_t i = threadIdx.x; i < 8 * 256; i += kScanThreads)
		tab[i] = a.d_tab[i];
	__syncthreads();
	if (warp >= L.nwarps)
		return;
	const int R = (int)L.ring; // multiple of 16, >= k + 16
	const uint32_t QL = a.qlane, nb = P.nbins;
	const WarpSmem sm = warp_smem(smem_raw + kTabBytes + (size_t)warp * fused_warp_bytes(L.ring, a.qlane, nb), L.ring, a.qlane, nb);
	uint2* planes = sm.planes;

	// the warp's open blocks: carried over from the previous launch unless a flush or a reset came between
	const uint32_t wg = blockIdx.x * kFusedWarps + warp;
	uint32_t* gs = P.gstate + ((size_t)a.ki * P.max_groups + wg) * gstate_row(nb);
	const uint32_t gen_f = P.ctl[CTL_FLUSHES] + 1u;
	const bool keep = wg < P.max_groups;
	{
		const bool resume = keep && gs[0] == P.epoch && gs[1] == gen_f;
		for (uint32_t b = lane; b < nb; b += 32) {
			sm.cur[b] = kVoid;
			sm.fill[b] = kBlkEntries;
			sm.curpos[b] = 0;
			if (resume) {
				sm.cur[b] = gs[2 + b];
				sm.fill[b] = gs[2 + nb + b];
				sm.curpos[b] = gs[2 + 2 * nb + b];
			}
		}
		__syncwarp();
	}

	uint64_t keep_pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_pol));
	const int k = (int)L.k;
	unsigned long long f1_local = 0;
	uint32_t n_flag = 0, n_defer = 0;
	bool exhausted = false; // the pool ran out while the sketch is not materialised: defer every further tile of this warp
	const uint32_t n_tiles = a.n_tiles, n_rec = a.n_rec, stride = a.stride;
	const uint32_t* __restrict__ words = a.words;
	for (uint32_t tile = warp * gridDim.x + blockIdx.x; tile < n_tiles; tile += gridDim.x * L.nwarps) {
		uint32_t p_from = 0;
		if (a.pass == 1) {
			const uint32_t ti = a.tile_info[tile];
			if (ti == kTileFlag || !(ti & kTileDefer))
				continue;
			p_from = ti & 0xFFFFu;
		}
		const uint32_t rb = tile * kTileRecs;
		const uint32_t nvalid = min(kTileRecs, n_rec - rb), last_rec = n_rec - 1u;
		int n = 0;
		bool mixed = false;
		for (bool first = true;; first = false) { // rounds: one unless a lane's queue overflowed
			uint4 v[32];
#pragma unroll
			for (int s = 0; s < 32; s++)
				v[s] = ldg_nc_v4(reinterpret_cast<const uint4*>(words + (uint64_t)min(rb + s * 32u + lane, last_rec) * stride), keep_pol);
			if (first) {
				uint32_t vmask = 0;
#pragma unroll
				for (int s = 0; s < 32; s++)
					vmask |= (s * 32u + lane < nvalid ? 1u : 0u) << s;
				const uint32_t len0 = __shfl_sync(0xFFFFFFFFu, v[0].x, 0);
				bool same = true;
#pragma unroll
				for (int s = 0; s < 32; s++)
					same = same && (v[s].x == len0);
				mixed = !__all_sync(0xFFFFFFFFu, same);
				n = (int)len0;
				bool flagged = false;
				if (mixed) {
					// records of different lengths: F1 (ntcard.cpp:155) is counted record by record
					unsigned long long c1 = 0;
					uint32_t mx = 0;
#pragma unroll
					for (int s = 0; s < 32; s++)
						if ((vmask >> s) & 1u) {
							const uint32_t ls = min(v[s].x, (stride - 1u) * 16u); // a corrupt length word counts as the record's capacity
							c1 += ls >= (uint32_t)k ? ls - (uint32_t)k + 1u : 0u;
							mx = max(mx, ls);
						}
					if (a.pass == 0)
						f1_local += c1;
					if (!L.mixed_ok || L.start_limit) {
						flagged = true; // the fallback kernel hashes this tile with the 64-bit recurrence
					} else {
#pragma unroll
						for (int d = 16; d > 0; d >>= 1)
							mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
						n = (int)mx;
					}
				}
				if (flagged) {
					if (lane == 0 && a.pass == 0) {
						a.tile_info[tile] = kTileFlag;
						n_flag++;
					}
					break;
				}
				n = min(n, (int)(stride - 1u) * 16); // a length word beyond the record's capacity (corrupt input)
				if (n >= k && L.start_limit)          // re-tiled pieces: only the first start_limit windows belong to this record
					n = min(n, (int)L.start_limit + k - 1);
				if (n >= k && !mixed && lane == 0 && a.pass == 0)
					f1_local += (unsigned long long)nvalid * (unsigned long long)(n - k + 1);
				if (n < k) {
					if (lane == 0 && a.pass == 0)
						a.tile_info[tile] = 0;
					break;
				}
				if (exhausted) {
					if (lane == 0) {
						a.tile_info[tile] = kTileDefer | p_from;
						n_defer++;
					}
					break;
				}
			}
			const uint32_t nwords = (uint32_t)(n + 15) >> 4;
			const uint32_t ngroups = (nwords + 1 + 3) / 4; // uint4 groups per record incl. the length word
			bs::State st;
#pragma unroll
			for (int j = 0; j < 31; j++) {
				st.F[j] = L.F0[j];
				st.R[j] = L.R0[j];
			}
			for (int j = 1; j <= k; j++) // the virtual positions -k..-1: no bases
				planes[(R - j) * 32 + lane] = make_uint2(0u, 0u);
			int cin = 0, cout = R - k;
			const int qlo = (int)p_from + k - 1; // first window end whose k-mer belongs to this round
			uint32_t cnt = 0, ovf_p = 0x7FFFFFFFu;
			uint16_t* const qbase = sm.queue + lane;
#pragma unroll 1
			for (uint32_t w = 0; w < nwords; w++) {
				const uint32_t g = (w + 1) >> 2, i = (w + 1) & 3u;
				uint32_t A[32];
				switch (i) {
				case 0:
#pragma unroll
					for (int s = 0; s < 32; s++) A[s] = v[s].x;
					break;
				case 1:
#pragma unroll
					for (int s = 0; s < 32; s++) A[s] = v[s].y;
					break;
				case 2:
#pragma unroll
					for (int s = 0; s < 32; s++) A[s] = v[s].z;
					break;
				default:
#pragma unroll
					for (int s = 0; s < 32; s++) A[s] = v[s].w;
					break;
				}
				if (i == 3 && g + 1 < ngroups) {
#pragma unroll
					for (int s = 0; s < 32; s++)
						v[s] = ldg_nc_v4(reinterpret_cast<const uint4*>(words + (uint64_t)min(rb + s * 32u + lane, last_rec) * stride) + (g + 1), keep_pol);
				}
				if (L.prefetch && first && w == nwords - 2) {
					// warm L2 with this warp's next tile one column before the end of this one (scan_kernel.cuh)
					const uint64_t nrb = (uint64_t)(tile + gridDim.x * L.nwarps) * kTileRecs;
					if (lane == 0 && nrb + kTileRecs <= n_rec && a.pass == 0) {
						const uint32_t bytes = kTileRecs * stride * 4u;
						asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(words + nrb * stride), "r"(bytes) : "memory");
					}
				}
				bs::transpose32(A);
				const int q0 = (int)(16u * w);
				{
					uint2* dst = planes + cin * 32 + lane; // a column never wraps: the ring size is a multiple of 16
#pragma unroll
					for (int j = 0; j < 16; j++)
						dst[j * 32] = make_uint2(A[2 * j], A[2 * j + 1]);
					if (cin == 0) { // mirror of slots 0..2 behind the ring
#pragma unroll
						for (int j = 0; j < 3; j++)
							dst[(R + j) * 32] = make_uint2(A[2 * j], A[2 * j + 1]);
					}
				}
				__syncwarp();
				const int nq = min(16, n - q0);
#pragma unroll 1
				for (int qb = q0; qb < q0 + nq; qb += kScanBlock) {
					int ob = cout + (qb - q0);
					ob = ob >= R ? ob - R : ob; // a block may start up to 3 slots before the end of the ring: mirror slots
					uint32_t m[kScanBlock];
					FusedBlock<KM, S, 0>::run(st, planes + lane + (cin + (qb - q0)) * 32, planes + lane + ob * 32, m);
					scan_rotate_home(st);
					if (qb + kScanBlock > qlo && !(a.dbg & 2u)) { // warp-uniform: some window of the block belongs to this round
#pragma unroll
						for (int u = 0; u < kScanBlock; u++) {
							const int q = qb + u;
							uint32_t x = (q >= qlo && q < n) ? m[u] : 0u;
							if (x) {
								const uint32_t p = (uint32_t)(q - (k - 1));
								if (cnt + (uint32_t)__popc(x) > QL) {
									ovf_p = min(ovf_p, p); // this lane's queue is full: the round ends before position p
								} else {
									const uint32_t pbits = p << 5;
									do {
										const uint32_t s = 31u - (uint32_t)__clz(x);
										x ^= 1u << s;
										qbase[cnt * 32u] = (uint16_t)(pbits | s);
										cnt++;
									} while (x);
								}
							}
						}
					}
				}
				cin = cin + 16 == R ? 0 : cin + 16;
				cout = cout + 16 >= R ? cout + 16 - R : cout + 16;
				__syncwarp();
			}
			uint32_t p_end = ovf_p;
#pragma unroll
			for (int d = 16; d > 0; d >>= 1)
				p_end = min(p_end, __shfl_xor_sync(0xFFFFFFFFu, p_end, d));
			const bool served = (a.dbg & 1u) ? true : fused_drain(a, tile, cnt, p_end, mixed, lane, warp);
			if (!served) {
				exhausted = true;
				if (lane == 0) {
					a.tile_info[tile] = kTileDefer | p_from;
					n_defer++;
				}
				break;
			}
			if (p_end == 0x7FFFFFFFu) {
				if (lane == 0)
					a.tile_info[tile] = 0;
				break;
			}
			p_from = p_end; // some lane's queue overflowed: scan the tile again, from the first position that did not fit
		}
	}
	// totKmer (ntcard.cpp:155), flagged and deferred tiles: warp-reduce, then one atomic per warp
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		f1_local += __shfl_xor_sync(0xFFFFFFFFu, f1_local, d);
	if (lane == 0) {
		if (f1_local)
			atomicAdd(a.f1_k, f1_local);
		if (n_flag)
			atomicAdd(P.ctl + CTL_NFLAG, n_flag);
		if (n_defer)
			atomicAdd(P.ctl + CTL_NDEFER, n_defer);
	}
	// keep the open blocks for the next launch
	if (keep) {
		__syncwarp();
		for (uint32_t b = lane; b < nb; b += 32) {
			gs[2 + b] = sm.cur[b];
			gs[2 + nb + b] = sm.fill[b];
			gs[2 + 2 * nb + b] = sm.curpos[b];
		}
		if (lane == 0) {
			gs[0] = P.epoch;
			gs[1] = gen_f;
		}
	}
}

template <int KM, int S>
cudaError_t launch_fused_one(const FusedArgs& a)
{
	auto kern = fused_kernel<KM, S>;
	static size_t smem_set[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64 || a.smem_bytes > smem_set[dev]) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem_bytes);
		if (e != cudaSuccess)
			return e;
		if (dev >= 0 && dev < 64)
			smem_set[dev] = a.smem_bytes;
	}
	kern<<<a.grid, kScanThreads, a.smem_bytes, a.stream>>>(a);
	return cudaGetLastError();
}

} // namespace pl
} // namespace ntc
