// bitslice_kernel.cuh -- the fast sm_100a sketch kernel: 32 records bit-sliced per lane.
//
// One warp owns a TILE of 1024 consecutive records of a uniform-stride batch (slot s of lane l is
// record tile*1024 + s*32 + l, so the 32 lanes of one load touch 32 neighbouring records).  Per tile
// and per k:
//   1. transpose: 128-bit loads of the packed bases, 32x32 bit transposes in registers, bit planes
//      (lo, hi per position, one 32-slot word each) to shared memory, [position][lane] -> conflict free;
//   2. scan: the bit-sliced upper-ring recurrence of bitslice_core.cuh over all positions, 31 positions
//      per unrolled body; the per-position "sampled" masks go to a 31-entry shared buffer;
//   3. hits: after each body the warp compacts the masks into a flat queue (popc + warp scan), then
//      every lane takes queued k-mers, re-hashes them in full 64 bits from the packed bases with
//      byte-indexed tables held in shared memory (8 x 128-bit lookups per 32 bases), and applies
//      ntComp (ntcard.cpp:132-145) -> RED.ADD into the uint32 sketch in HBM.
// Tiles whose records are not all the same length (or the last partial tile) take the general
// 64-bit path in place (process_piece_k), so any uniform-stride batch is handled exactly.
//
// No tensor cores: there is no dense contraction here; the kernel is bound by the 16-lane integer
// ALU pipe (LOP3) and by the random RED traffic of the sketch -- see DESIGN.md.
#pragma once
#include <cuda_runtime.h>

#include "bitslice_core.cuh"
#include "bitslice_launch.h"
#include "sketch_common.cuh"

namespace ntc {
namespace bs {

constexpr int kWarpsMax = 4;
constexpr uint32_t kTileRecs = 1024;
constexpr uint32_t kQueue = 1024; // flat hit queue entries per warp per body (expected 31*1024/64 = 496 at s=7)

struct WarpCtx {
	const uint32_t* __restrict__ words;
	uint32_t stride;
	uint32_t rb;        // first record of the tile
	uint32_t k, rBits, sBits;
	uint32_t nwords;    // base words per record actually holding bases
	const uint4* tab;   // shared: [8][256] {FB.lo, FB.hi, RB.lo, RB.hi}
	uint32_t* __restrict__ ctr_k;
};

// Full canonical hash of the k-mer starting at base p of a record (b = its base words).
// k = t + 32*M: the t head bases one at a time (NTF64/NTR64 base forms, nthash.hpp:220-239), then M
// blocks of 32 bases through the byte tables:
//   fh = srol^32(fh) ^ FB_m,  FB = XOR_i srol^(31-i) seed[c_i]     rh ^= srol^(t+32m) RB_m,  RB = XOR_i srol^i seed[3-c_i]
__device__ __forceinline__ void hash_kmer(const WarpCtx& c, const uint32_t* __restrict__ b, uint32_t p, uint64_t& fh, uint64_t& rh)
{
	const uint32_t t = c.k & 31u, M = c.k >> 5;
	fh = 0;
	rh = 0;
	for (uint32_t i = 0; i < t; i++) {
		const uint32_t code = base_at(b, p + i);
		fh = srol(fh) ^ seed_of(code);
		rh ^= srol_n(seed_of(3u - code), i);
	}
	for (uint32_t m = 0; m < M; m++) {
		const uint32_t o = p + t + 32u * m, wi = o >> 4, sh = (o & 15u) * 2u;
		const uint32_t x0 = __ldg(b + wi), x1 = __ldg(b + wi + 1), x2 = __ldg(b + min(wi + 2, c.nwords - 1));
		const uint32_t w0 = __funnelshift_r(x0, x1, sh), w1 = __funnelshift_r(x1, x2, sh);
		uint32_t f0 = 0, f1 = 0, r0 = 0, r1 = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const uint32_t byte = ((j < 4 ? w0 : w1) >> (8 * (j & 3))) & 0xFFu;
			const uint4 e = c.tab[j * 256 + byte];
			f0 ^= e.x;
			f1 ^= e.y;
			r0 ^= e.z;
			r1 ^= e.w;
		}
		fh = srol_n(fh, 32) ^ (((uint64_t)f1 << 32) | f0);
		rh ^= srol_n(((uint64_t)r1 << 32) | r0, t + 32u * m);
	}
}

// entry: slot (5 bits) | lane (5 bits) << 5 | position-in-body << 10
__device__ __forceinline__ void process_hit(const WarpCtx& c, uint32_t e, uint32_t q0)
{
	const uint32_t s = e & 31u, ln = (e >> 5) & 31u, tq = e >> 10;
	const uint32_t rec = c.rb + s * 32u + ln;
	const uint32_t p = q0 + tq + 1u - c.k;
	const uint32_t* b = c.words + (uint64_t)rec * c.stride + 1;
	uint64_t fh, rh;
	hash_kmer(c, b, p, fh, rh);
	sample_and_count(rh < fh ? rh : fh, c.ctr_k, c.rBits, c.sBits);
}

// Compact the masks of one body (nq positions) into the queue and hash the hits.
static __device__ __noinline__ void drain_body(const WarpCtx& c, const uint32_t* __restrict__ hw, uint32_t* __restrict__ queue, uint32_t nq,
    uint32_t q0, uint32_t lane)
{
	uint32_t cnt = 0;
	for (uint32_t tq = 0; tq < nq; tq++)
		cnt += __popc(hw[tq * 32 + lane]);
	uint32_t inc = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
		if ((int)lane >= d)
			inc += v;
	}
	const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
	if (total == 0)
		return;
	uint32_t off = inc - cnt;
	for (uint32_t tq = 0; tq < nq; tq++) {
		uint32_t w = hw[tq * 32 + lane];
		while (w) {
			const uint32_t s = __ffs(w) - 1;
			w &= w - 1;
			const uint32_t e = s | (lane << 5) | (tq << 10);
			if (off < kQueue)
				queue[off] = e;
			else
				process_hit(c, e, q0); // queue overflow (heavily skewed data): slow but exact
			off++;
		}
	}
	__syncwarp();
	const uint32_t lim = min(total, kQueue);
	for (uint32_t i = lane; i < lim; i += 32)
		process_hit(c, queue[i], q0);
	__syncwarp();
}

template <int KM, int S, int TQ> struct DevBody {
	static __device__ __forceinline__ void run(State& st, const uint2* __restrict__ pl, uint32_t* __restrict__ hw, int q0, int n, int k)
	{
		const int q = q0 + TQ;
		if (q < n) {
			const uint2 in = pl[(q + 1) * 32];
			int oq = q - k + 1; // plane slot of the leaving base (position q-k), slot 0 = all-zero sentinel
			oq = oq < 0 ? 0 : oq;
			const uint2 out = pl[oq * 32];
			step<KM, TQ>(st, in.x, in.y, out.x, out.y);
			uint32_t m = 0;
			if (q >= k - 1)
				m = sampled_mask<TQ, S>(st);
			hw[TQ * 32] = m;
			DevBody<KM, S, TQ + 1>::run(st, pl, hw, q0, n, k);
		}
	}
};
template <int KM, int S> struct DevBody<KM, S, 31> {
	static __device__ __forceinline__ void run(State&, const uint2* __restrict__, uint32_t* __restrict__, int, int, int) {}
};

template <int KM, int S>
__global__ void __launch_bounds__(kWarpsMax * 32, 1) bitslice_kernel(const uint32_t* __restrict__ words, uint32_t stride, uint32_t n_rec,
    BsLaunch L, const uint4* __restrict__ g_tab, const DevParams* __restrict__ P, uint32_t* __restrict__ ctr_k,
    unsigned long long* __restrict__ f1_k)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	uint4* tab = reinterpret_cast<uint4*>(smem_raw);
	for (uint32_t i = threadIdx.x; i < 8 * 256; i += blockDim.x)
		tab[i] = g_tab[i];
	const uint32_t plane_bytes = (1u + L.pos_cap) * 256u; // [slot 0 = zeros][position][lane] uint2
	unsigned char* wbase = smem_raw + 8 * 256 * 16 + warp * (plane_bytes + 31 * 128 + kQueue * 4);
	uint2* planes = reinterpret_cast<uint2*>(wbase);
	uint32_t* hwbuf = reinterpret_cast<uint32_t*>(wbase + plane_bytes);
	uint32_t* queue = hwbuf + 31 * 32;
	planes[lane] = make_uint2(0u, 0u);
	__syncthreads();

	WarpCtx c;
	c.words = words;
	c.stride = stride;
	c.k = L.k;
	c.rBits = L.rBits;
	c.sBits = S;
	c.tab = tab;
	c.ctr_k = ctr_k;
	const int k = (int)L.k;
	unsigned long long f1_local = 0;
	const uint32_t n_tiles = (n_rec + kTileRecs - 1) / kTileRecs;
	for (uint32_t tile = blockIdx.x * nwarps + warp; tile < n_tiles; tile += gridDim.x * nwarps) {
		const uint32_t rb = tile * kTileRecs;
		c.rb = rb;
		// ---- record lengths; is the tile uniform? ----
		uint32_t len0 = 0;
		bool uniform = true;
		{
			const uint32_t first_len = rb < n_rec ? __ldg(words + (uint64_t)rb * stride) : 0u;
#pragma unroll 4
			for (uint32_t s = 0; s < 32; s++) {
				const uint32_t rec = rb + s * 32u + lane;
				const uint32_t len = rec < n_rec ? __ldg(words + (uint64_t)rec * stride) : 0xFFFFFFFFu;
				uniform = uniform && (len == first_len);
			}
			uniform = __all_sync(0xFFFFFFFFu, uniform);
			len0 = first_len;
		}
		if (!uniform || (len0 != 0xFFFFFFFFu && len0 > L.pos_cap)) {
			// general path, in place: each lane walks its own records with the 64-bit recurrence
			const KTab& T = P->tab[L.ki];
			uint32_t cnt = 0;
			for (uint32_t s = 0; s < 32; s++) {
				const uint32_t rec = rb + s * 32u + lane;
				if (rec < n_rec) {
					const uint32_t* r = words + (uint64_t)rec * stride;
					cnt += process_piece_k(r + 1, __ldg(r), 0, L.k, T, ctr_k, L.rBits, S);
				}
			}
			f1_local += cnt;
			continue;
		}
		const int n = (int)len0;
		if (n < k)
			continue;
		c.nwords = (uint32_t)(n + 15) >> 4;
		// ---- 1. transpose packed bases into bit planes ----
		{
			const uint32_t ngroups = (c.nwords + 1 + 3) / 4; // uint4 groups per record incl. the length word
			for (uint32_t g = 0; g < ngroups; g++) {
				uint4 v[32];
#pragma unroll
				for (int s = 0; s < 32; s++)
					v[s] = __ldg(reinterpret_cast<const uint4*>(words + (uint64_t)(rb + s * 32u + lane) * stride) + g);
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const int w = (int)(g * 4) + i - 1; // base word index
					if (w < 0 || w >= (int)c.nwords)
						continue;
					uint32_t A[32];
#pragma unroll
					for (int s = 0; s < 32; s++)
						A[s] = i == 0 ? v[s].x : i == 1 ? v[s].y : i == 2 ? v[s].z : v[s].w;
					transpose32(A);
#pragma unroll
					for (int j = 0; j < 16; j++)
						if (16 * w + j < n)
							planes[(1 + 16 * w + j) * 32 + lane] = make_uint2(A[2 * j], A[2 * j + 1]);
				}
			}
		}
		__syncwarp();
		// ---- 2./3. scan + hits ----
		State st;
#pragma unroll
		for (int j = 0; j < 31; j++) {
			st.F[j] = L.F0[j];
			st.R[j] = L.R0[j];
		}
		for (int q0 = 0; q0 < n; q0 += 31) {
			DevBody<KM, S, 0>::run(st, planes + lane, hwbuf + lane, q0, n, k);
			__syncwarp();
			const int nq = min(31, n - q0);
			if (q0 + nq >= k) // some position of this body ends a full window
				drain_body(c, hwbuf, queue, (uint32_t)nq, (uint32_t)q0, lane);
		}
		if (lane == 0)
			f1_local += (unsigned long long)kTileRecs * (unsigned long long)(n - k + 1);
	}
	// totKmer (ntcard.cpp:155): warp-reduce then one atomic per warp
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		f1_local += __shfl_xor_sync(0xFFFFFFFFu, f1_local, d);
	if (lane == 0 && f1_local)
		atomicAdd(f1_k, f1_local);
}

template <int KM, int S>
cudaError_t launch_one(const BsArgs& a)
{
	auto kern = bitslice_kernel<KM, S>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem_bytes);
	if (e != cudaSuccess)
		return e;
	kern<<<a.grid, a.warps * 32, a.smem_bytes, a.stream>>>(a.words, a.stride, a.n_rec, a.L, a.d_tab, a.d_params, a.ctr_k, a.f1_k);
	return cudaGetLastError();
}

} // namespace bs
} // namespace ntc
