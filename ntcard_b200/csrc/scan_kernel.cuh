// scan_kernel.cuh -- stage S of the sketch pipeline: the bit-sliced sampling filter (bitslice_core.cuh) over a
// uniform-stride batch.  8 warps per CTA (2 per SM sub-partition, so one warp's loads and shared-memory round
// trips hide behind the other's LOP3 stream), one CTA per SM, up to 255 registers per thread.
//
// A warp owns a TILE of 1024 consecutive records (slot s of lane l is record tile*1024 + s*32 + l).  Per column
// of 16 positions: 128-bit loads of the packed bases (4 columns per load, requested one column ahead), a 32x32
// bit transpose in registers, bit planes to a shared-memory ring, then the scan: 62 LOP3 for the two 31-bit upper
// rings + ~30 for the word preparation and the bit-sliced ntComp test per position, for 32 k-mers per lane.
// Output: one mask word per (tile, k-mer position, lane) -- bit s set = the k-mer of slot s ending here is
// sampled by ntComp (ntcard.cpp:132-145) -- written straight to HBM with coalesced 128-byte stores; the hit
// kernel (hit_kernels.cu) turns the set bits into sketch increments.  Also: F1 (ntcard.cpp:155), the candidate
// count of the batch, and per-tile info (k-mer positions per record).  Tiles whose records differ in length are scanned
// at the longest length (the hit kernel drops what lies past a record's end), or flagged for the fallback kernel when
// the all-A k-mer of the padding would itself be sampled.
//
// No tensor cores: there is no dense contraction; the kernel is bound by the 16-lane integer ALU pipe.
#pragma once
#include <cuda_runtime.h>

#include "bitslice_core.cuh"
#include "pipeline.h"
#include "sketch_common.cuh"

namespace ntc {
namespace pl {

constexpr int kScanBlock = 4;      // scan positions per unrolled block
constexpr int kScanThreads = 256;  // 8 warps

// 128-bit loads of the packed reads with an L2 evict_last hint: a record's three 16-byte groups are read tens of
// microseconds apart, and without the hint the 32-byte sectors they share leave L2 in between (2.4x the input
// bytes came from DRAM).
__device__ __forceinline__ uint4 ldg_nc_v4(const uint4* p, uint64_t policy)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	             : "l"(p), "l"(policy));
	return r;
}

// Branch free on purpose: positions past the end of the read run on whatever the planes hold and store nothing.
// Planes live in a ring of R positions (R = k + 16 rounded up to 16; slot = position mod R) followed by 3 slots that mirror slots
// 0..2, so the kScanBlock slots of a block can be addressed from one base without wrapping.  The k "virtual"
// positions before the read (window not full yet) are the k slots before slot 0, zeroed at the start of a tile.
// pin / pout / mb: plane slot of the block's first entering / leaving position, mask row of its first position.
// MODE 0: ntComp's sampling predicate with sBits = S; MODE 1: nthll pre-filter, top S bits of the canonical hash all zero;
// MODE 2: nthll pre-filter at the run-time level lev (bitslice_core.cuh zero_top_mask_lev)
template <int KM, int S, int U, int MODE> struct ScanBlock {
	static __device__ __forceinline__ void run(bs::State& st, const uint2* __restrict__ pin, const uint2* __restrict__ pout,
	    uint32_t* __restrict__ mb, int qb, int k, int n, uint32_t& cand, uint32_t lev)
	{
		const uint2 in = pin[U * 32];
		const uint2 out = pout[U * 32];
		bs::step<KM, U>(st, in.x, in.y, out.x, out.y);
		const uint32_t m = MODE == 2 ? bs::zero_top_mask_lev<U>(st, lev) : MODE ? bs::zero_top_mask<U, S>(st) : bs::sampled_mask<U, S>(st); // slots past the end of the batch are dropped by the hit kernel
		const int q = qb + U;
		if (q >= k - 1 && q < n) {
			__stcs(mb + U * 32, m); // streaming store: read once by the hit kernel
			cand += __popc(m);
		}
		ScanBlock<KM, S, U + 1, MODE>::run(st, pin, pout, mb, qb, k, n, cand, lev);
	}
};
template <int KM, int S, int MODE> struct ScanBlock<KM, S, kScanBlock, MODE> {
	static __device__ __forceinline__ void run(bs::State&, const uint2* __restrict__, const uint2* __restrict__, uint32_t* __restrict__, int, int, int,
	    uint32_t&, uint32_t)
	{
	}
};

// After kScanBlock steps logical ring bit r sits in physical F[r - kScanBlock] / R[r + kScanBlock]: move it home
// (the compiler absorbs the moves into the destination registers of the block's last position).
__device__ __forceinline__ void scan_rotate_home(bs::State& st)
{
	bs::State t;
#pragma unroll
	for (int j = 0; j < 31; j++) {
		t.F[j] = st.F[bs::mod31(j - kScanBlock)];
		t.R[j] = st.R[bs::mod31(j + kScanBlock)];
	}
	st = t;
}

// mask rows [r0, npos_max) of a tile hold no k-mer: all-zero words (the hit kernel reads every row of every tile)
__device__ __forceinline__ void zero_rows(uint32_t* __restrict__ masks, uint32_t tile, uint32_t r0, uint32_t npos_max, uint32_t lane)
{
	uint32_t* p = masks + ((size_t)tile * npos_max) * 32 + lane;
	for (uint32_t r = r0; r < npos_max; r++)
		__stcs(p + (size_t)r * 32, 0u);
}

template <int KM, int S, int MODE>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const uint32_t* __restrict__ words, uint32_t stride, uint32_t n_rec, ScanLaunch L,
    uint32_t* __restrict__ masks, uint32_t* __restrict__ tile_info, unsigned long long* __restrict__ f1_k,
    unsigned long long* __restrict__ cand_out, uint32_t* __restrict__ ctl, const uint32_t* __restrict__ hll_min)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	if (warp >= L.nwarps)
		return;
	// MODE 2: the filter follows the smallest register as the device has it now (hll_min_kernel ran on this stream after the
	// previous chunk): top T bits zero is necessary for a k-mer to raise a register once every register is >= T - 1
	uint32_t lev = 0;
	if (MODE == 2) {
		const uint32_t mn = __ldcg(hll_min);
		lev = mn >= 12 ? 5u : mn >= 10 ? 4u : mn >= 8 ? 3u : mn >= 6 ? 2u : mn >= 4 ? 1u : 0u;
	}
	const int R = (int)L.ring; // multiple of 16, >= k + 16
	uint2* planes = reinterpret_cast<uint2*>(smem_raw + (size_t)warp * (L.ring + 3u) * 256u); // [position mod ring (+3 mirror slots)][lane]

	uint64_t keep;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
	const int k = (int)L.k;
	unsigned long long f1_local = 0, cand_local = 0;
	uint32_t n_flag = 0;
	const uint32_t n_tiles = (n_rec + kTileRecs - 1) / kTileRecs;
	// tile = round * (grid * nwarps) + warp * grid + cta: the tiles of the last, partial round go to warp 0 (then 1, ...)
	// of EVERY SM instead of to all warps of a few SMs
	for (uint32_t tile = warp * gridDim.x + blockIdx.x; tile < n_tiles; tile += gridDim.x * L.nwarps) {
		const uint32_t rb = tile * kTileRecs;
		// The last tile may be partial: slots past the end re-read the batch's last record (every load stays in
		// bounds, the uniformity test is unaffected); the hit kernel drops their candidates (record index >= n_rec).
		const uint32_t nvalid = min(kTileRecs, n_rec - rb), last_rec = n_rec - 1u;
		uint32_t vmask = 0;
		uint4 v[32];
#pragma unroll
		for (int s = 0; s < 32; s++) {
			v[s] = ldg_nc_v4(reinterpret_cast<const uint4*>(words + (uint64_t)min(rb + s * 32u + lane, last_rec) * stride), keep);
			vmask |= (s * 32u + lane < nvalid ? 1u : 0u) << s;
		}
		const uint32_t len0 = __shfl_sync(0xFFFFFFFFu, v[0].x, 0);
		bool same = true;
#pragma unroll
		for (int s = 0; s < 32; s++)
			same = same && (v[s].x == len0);
		const bool mixed = !__all_sync(0xFFFFFFFFu, same);
		int n = (int)len0;
		if (mixed) {
			// records of different lengths: their k-mers are counted here, record by record (F1, and as an upper bound of
			// the tile's sketch increments)
			unsigned long long cnt = 0;
			uint32_t mx = 0;
#pragma unroll
			for (int s = 0; s < 32; s++)
				if ((vmask >> s) & 1u) {
					const uint32_t ls = min(v[s].x, (stride - 1u) * 16u); // a corrupt length word counts as the record's capacity
					cnt += ls >= (uint32_t)k ? ls - (uint32_t)k + 1u : 0u;
					mx = max(mx, ls);
				}
			f1_local += cnt;
			if (!L.mixed_ok || L.start_limit) {
				// the fallback kernel hashes this tile with the 64-bit recurrence
				cand_local += cnt;
				if (lane == 0) {
					tile_info[tile] = kTileFlag;
					n_flag++;
				}
				zero_rows(masks, tile, 0, L.npos_max, lane);
				continue;
			}
			// Scan the tile as if every record had the longest length: positions past a shorter record's end run on its
			// padding and may mark candidates, which the hit kernel drops (start + k > the record's length).  Only legal
			// when the all-A k-mer of the zero padding is not itself sampled (mixed_ok, decided on the host per k).
#pragma unroll
			for (int d = 16; d > 0; d >>= 1)
				mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
			n = (int)mx;
		}
		n = min(n, (int)(stride - 1u) * 16); // a length word beyond the record's capacity (corrupt input) must not leave the tile's mask rows
		if (n >= k && L.start_limit) // re-tiled pieces: only the first start_limit windows belong to this record
			n = min(n, (int)L.start_limit + k - 1);
		if (n < k) {
			if (lane == 0)
				tile_info[tile] = 0;
			zero_rows(masks, tile, 0, L.npos_max, lane);
			continue;
		}
		zero_rows(masks, tile, (uint32_t)(n - k + 1), L.npos_max, lane);
		if (lane == 0) {
			tile_info[tile] = (uint32_t)(n - k + 1);
			if (mixed) // counted record by record above: cancel the uniform formula at the end of the tile
				f1_local -= (unsigned long long)nvalid * (unsigned long long)(n - k + 1);
		}
		const uint32_t nwords = (uint32_t)(n + 15) >> 4;
		const uint32_t ngroups = (nwords + 1 + 3) / 4; // uint4 groups per record incl. the length word
		// mask row of the block's first position: a running pointer (not recomputed from tile / lane per store)
		uint32_t* mptr = masks + ((long long)tile * L.npos_max - (k - 1)) * 32 + lane;
		bs::State st;
#pragma unroll
		for (int j = 0; j < 31; j++) {
			st.F[j] = L.F0[j];
			st.R[j] = L.R0[j];
		}
		uint32_t cand = 0;
		for (int j = 1; j <= k; j++) // the virtual positions -k..-1: no bases
			planes[(R - j) * 32 + lane] = make_uint2(0u, 0u);
		int cin = 0, cout = R - k; // ring slots of the column's first entering position and of its first leaving position
		// One column = one packed word of every record = 16 positions: transpose it into the plane ring, scan it.
		// The next 16 bytes of every record are requested right after the last word of the current 16 bytes has
		// been taken out of v[], so the loads fly during a whole column's scan.
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
					v[s] = ldg_nc_v4(reinterpret_cast<const uint4*>(words + (uint64_t)min(rb + s * 32u + lane, last_rec) * stride) + (g + 1), keep);
			}
			if (L.prefetch && w == (L.prefetch == 1 ? nwords / 2 : nwords - 2)) {
				// warm L2 with this warp's next tile (bulk async prefetch) -- late: one column before the end of this tile.  Half way
				// through (mode 1) doubles the tiles resident in L2 (2 x 57 MB for 1184 warps) and the 32-byte sectors that a record's
				// 16-byte groups share were evicted between their uses: 876 MB of DRAM reads per 10 M reads instead of 492 MB
				// (profiles/r01_scan_l2_experiment.txt)
				const uint64_t nrb = (uint64_t)(tile + gridDim.x * L.nwarps) * kTileRecs;
				if (lane == 0 && nrb + kTileRecs <= n_rec) {
					const uint32_t bytes = kTileRecs * stride * 4u;
					asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(words + nrb * stride), "r"(bytes) : "memory");
				}
			}
			bs::transpose32(A);
			const int q0 = (int)(16u * w);
			{
				const int c0 = cin;
				uint2* dst = planes + c0 * 32 + lane; // a column never wraps: the ring size is a multiple of 16
#pragma unroll
				for (int j = 0; j < 16; j++)
					dst[j * 32] = make_uint2(A[2 * j], A[2 * j + 1]);
				if (c0 == 0) { // mirror of slots 0..2 behind the ring
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
				ScanBlock<KM, S, 0, MODE>::run(st, planes + lane + (cin + (qb - q0)) * 32, planes + lane + ob * 32, mptr, qb, k, n, cand, lev);
				scan_rotate_home(st);
				mptr += kScanBlock * 32;
			}
			cin = cin + 16 == R ? 0 : cin + 16;
			cout = cout + 16 >= R ? cout + 16 - R : cout + 16;
			__syncwarp();
		}
		cand_local += cand;
		if (lane == 0)
			f1_local += (unsigned long long)nvalid * (unsigned long long)(n - k + 1);
	}
	// totKmer (ntcard.cpp:155), candidate count, flagged tiles: warp-reduce, then one atomic per warp
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		f1_local += __shfl_xor_sync(0xFFFFFFFFu, f1_local, d);
		cand_local += __shfl_xor_sync(0xFFFFFFFFu, cand_local, d);
	}
	if (lane == 0) {
		if (f1_local)
			atomicAdd(f1_k, f1_local);
		if (cand_local)
			atomicAdd(cand_out, cand_local);
		if (n_flag)
			atomicAdd(ctl + CTL_NFLAG, n_flag);
	}
}

template <int KM, int S, int MODE = 0>
cudaError_t launch_scan_one(const ScanArgs& a)
{
	auto kern = scan_kernel<KM, S, MODE>;
	// the opt-in shared-memory limit is a property of (function, device): set it when it has to grow, not on every launch
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
	kern<<<a.grid, kScanThreads, a.smem_bytes, a.stream>>>(a.words, a.stride, a.n_rec, a.L, a.masks, a.tile_info, a.f1_k, a.cand, a.ctl, a.hll_min);
	return cudaGetLastError();
}

} // namespace pl
} // namespace ntc
