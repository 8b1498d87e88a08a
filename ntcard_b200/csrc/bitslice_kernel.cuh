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

constexpr int kBlock = 4;    // scan positions per unrolled block (see DevBlock)
constexpr int kBodyPos = 16; // positions per hand-off to a hit warp (4 blocks)
constexpr int kPairsMax = 4;       // scan warps per CTA; each has one partner "hit" warp on the same SM sub-partition
constexpr uint32_t kTileRecs = 1024;

struct WarpCtx {
	const uint32_t* __restrict__ words;
	uint32_t stride;
	uint32_t rb;        // first record of the tile
	uint32_t k, rBits;
	uint32_t nwords;    // base words per record actually holding bases
	const uint4* tab;   // shared: [8][256] {FB.lo, FB.hi, RB.lo, RB.hi}
	uint64_t rot_a, rot_b; // byte m: (t + 32m) % 31 and % 33 (rotation of RB_m)
	uint32_t* __restrict__ ctr_k;
	uint64_t red_policy;   // L2 evict_first cache policy for the sketch increments
	uint64_t keep_policy;  // L2 evict_last cache policy for the packed reads
	uint32_t nored;
};

// ---- mbarrier helpers (shared::cta) -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "MBAR_WAIT:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra MBAR_DONE;\n\t"
	    "bra MBAR_WAIT;\n\t"
	    "MBAR_DONE:\n\t"
	    "}" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

// Consumer-side wait: poll, and sleep between polls -- an idle hit warp must not eat the issue slots of the scan
// warp it shares a sub-partition with (a bare try_wait loop there was a third of all issued instructions).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity)
{
	uint32_t done = 0;
	for (;;) {
		asm volatile(
		    "{\n\t"
		    ".reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t"
		    "}"
		    : "=r"(done)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
		if (done)
			break;
		__nanosleep(128);
	}
}

// Named barriers (bar.sync / bar.arrive) for the producer -> consumer direction: a waiting hit warp is
// descheduled by the hardware and costs no issue slots (an mbarrier try_wait loop does: a third of all
// issued instructions were idle hit warps polling).
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads)
{
	asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 128-bit / 32-bit loads of the packed reads with an L2 evict_last hint: the hit path comes back to these lines
// ~100 us later, after gigabytes of sketch sectors have streamed through L2.
__device__ __forceinline__ uint4 ldg_keep_v4(const uint4* p, uint64_t policy)
{
	uint4 r;
	asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	             : "l"(p), "l"(policy));
	return r;
}
__device__ __forceinline__ uint32_t ldg_keep_u32(const uint32_t* p, uint64_t policy)
{
	uint32_t r;
	asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy));
	return r;
}

// srol applied n times, for n given as (a, b) = (n % 31, n % 33): rotate the upper ring by a, the lower by b.
__device__ __forceinline__ uint64_t srol_ab(uint64_t v, uint32_t a, uint32_t b)
{
	uint32_t hi = (uint32_t)(v >> 33);
	uint64_t lo = v & 0x1FFFFFFFFull;
	hi = ((hi << a) | (hi >> (31u - a))) & 0x7FFFFFFFu;
	lo = ((lo << b) | (lo >> (33u - b))) & 0x1FFFFFFFFull;
	return ((uint64_t)hi << 33) | lo;
}

struct HitLoad { // a queued k-mer with the packed words of its first 32-base block in flight
	uint32_t p;
	const uint32_t* __restrict__ b;
	uint32_t x0, x1, x2;
};

// entry: slot (5 bits) | lane (5 bits) << 5 | position-in-body << 10
__device__ __forceinline__ HitLoad hit_issue(const WarpCtx& c, uint32_t e, uint32_t q0)
{
	HitLoad h;
	const uint32_t s = e & 31u, ln = (e >> 5) & 31u, tq = e >> 10;
	const uint32_t rec = c.rb + s * 32u + ln;
	h.p = q0 + tq + 1u - c.k;
	h.b = c.words + (uint64_t)rec * c.stride + 1;
	const uint32_t wi = (h.p + (c.k & 31u)) >> 4;
	h.x0 = ldg_keep_u32(h.b + min(wi, c.nwords - 1), c.keep_policy);
	h.x1 = ldg_keep_u32(h.b + min(wi + 1, c.nwords - 1), c.keep_policy);
	h.x2 = ldg_keep_u32(h.b + min(wi + 2, c.nwords - 1), c.keep_policy);
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

// Full canonical hash of a queued k-mer and ntComp (ntcard.cpp:132-145).  k = t + 32*M: the t head bases
// one at a time (NTF64/NTR64 base forms, nthash.hpp:220-239), then M blocks of 32 bases:
//   fh = srol^32(fh) ^ FB_m        rh ^= srol^(t+32m) RB_m
template <int S> __device__ __forceinline__ void hit_finish(const WarpCtx& c, const HitLoad& h)
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
				x2 = __ldg(h.b + min(wi + 2, c.nwords - 1));
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
	const bool t0 = (hh >> (31 - S)) == 1u;
	const bool t1 = (hh >> (32 - S)) == ((1u << (S - 1)) - 1u);
	if ((t0 || t1) && !c.nored) {
		const uint32_t idx = ((t1 ? 1u : 0u) << c.rBits) | (hl & ((1u << c.rBits) - 1u));
		// fire-and-forget increment; the sketch sector is streaming data (one touch), so let it leave L2 first and
		// keep the packed reads of the tiles in flight resident (the hit path re-reads them)
		asm volatile("red.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(c.ctr_k + idx), "r"(1u), "l"(c.red_policy) : "memory");
	}
}

constexpr int kHitBatch = 4;  // k-mers per lane whose loads are in flight together

// queue-overflow path, kept out of line
template <int S> static __device__ __noinline__ void hit_slow(const WarpCtx& c, uint32_t e, uint32_t q0)
{
	hit_finish<S>(c, hit_issue(c, e, q0));
}

// A hit warp's work for one hand-off unit (<= 16 positions): count the sampled k-mers (popc + warp prefix
// sum), write them as a flat queue, then hash them kHitBatch per lane at a time.  Deliberately compact
// code (loops, not unrolled): the instruction cache is shared with the scan warp on the same sub-partition,
// and an unrolled variant of this function made the scan warp starve on instruction fetch.
template <int S>
__device__ __forceinline__ void drain_body(const WarpCtx& c, const uint32_t* __restrict__ hw, uint32_t* __restrict__ queue, uint32_t nq,
    uint32_t q0, uint32_t lane)
{
	uint32_t cnt = 0;
#pragma unroll 4
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
			if (off < kQueueCap)
				queue[off] = e;
			else
				hit_slow<S>(c, e, q0); // queue overflow (skewed data): slow but exact
			off++;
		}
	}
	__syncwarp();
	const uint32_t lim = min(total, (uint32_t)kQueueCap);
#pragma unroll 1
	for (uint32_t base = 0; base < lim; base += 32 * kHitBatch) {
		HitLoad h[kHitBatch];
#pragma unroll
		for (int u = 0; u < kHitBatch; u++) {
			const uint32_t i = base + u * 32 + lane;
			if (i < lim)
				h[u] = hit_issue(c, queue[i], q0);
		}
#pragma unroll
		for (int u = 0; u < kHitBatch; u++) {
			const uint32_t i = base + u * 32 + lane;
			if (i < lim)
				hit_finish<S>(c, h[u]);
		}
	}
	__syncwarp();
}

// The scan is unrolled over kBlock positions only: after a block the 31+31 state registers are rotated
// back into their home frame (register moves, which issue on the FMA pipe as IMAD.MOV and leave the
// LOP3 pipe alone), so one ~14 KB block of code serves every position.  Unrolling the full ring
// period (31) needs no moves but is a 52 KB loop, and four scan warps in different phases of it
// starved on instruction fetch (43 % of their stall samples were no_instruction).

template <int KM, int S, int U> struct DevBlock {
	// Branch free on purpose: positions past the end of the read run on whatever the planes hold (their
	// masks are never consumed: the hand-off carries the number of valid positions).
	// Planes live in a ring of rmask+1 positions (slot = position mod ring size); slot rmask+1 is all zero and
	// stands for the virtual bases before the read (window not full yet).
	static __device__ __forceinline__ void run(State& st, const uint2* __restrict__ pl, uint32_t* __restrict__ hw, int qb, int k, int rmask, uint32_t vmask)
	{
		const int q = qb + U;
		const uint2 in = pl[(q & rmask) * 32];
		const int oq = q - k;
		const uint2 out = pl[(oq < 0 ? rmask + 1 : (oq & rmask)) * 32];
		step<KM, U>(st, in.x, in.y, out.x, out.y);
		const uint32_t m = sampled_mask<U, S>(st);
		hw[U * 32] = q >= k - 1 ? (m & vmask) : 0u;
		DevBlock<KM, S, U + 1>::run(st, pl, hw, qb, k, rmask, vmask);
	}
};
template <int KM, int S> struct DevBlock<KM, S, kBlock> {
	static __device__ __forceinline__ void run(State&, const uint2* __restrict__, uint32_t* __restrict__, int, int, int, uint32_t) {}
};

// After kBlock steps logical ring bit r sits in physical F[r - kBlock] / R[r + kBlock]: move it home.
__device__ __forceinline__ void rotate_home(State& st)
{
	State t;
#pragma unroll
	for (int j = 0; j < 31; j++) {
		t.F[j] = st.F[mod31(j - kBlock)];
		t.R[j] = st.R[mod31(j + kBlock)];
	}
	st = t;
}

// Per scan warp: hand-off of one body's masks to the partner hit warp.
struct BodyDesc {
	uint32_t rb, q0, nq, nwords; // nq == 0: no more work
};

constexpr int kHitWarps = 3;                        // hit warps per scan warp (hand-off unit i goes to hit warp i % 3)
constexpr int kThreads = 128 * (1 + kHitWarps);     // warpgroup 0: 4 scan warps; warpgroups 1..3: their hit warps
constexpr int kScanRegs = 240, kHitRegs = 88;       // 128*240 + 384*88 = 64512 of the 65536 registers (asking for all of them never completes)

template <int KM, int S>
__global__ void __launch_bounds__(kThreads, 1) bitslice_kernel(const uint32_t* __restrict__ words, uint32_t stride, uint32_t n_rec,
    BsLaunch L, const uint4* __restrict__ g_tab, const DevParams* __restrict__ P, uint32_t* __restrict__ ctr_k,
    unsigned long long* __restrict__ f1_k)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t role = warp >> 2;  // 0 scan, 1..kHitWarps hit
	const uint32_t pair = warp & 3u;
	const uint32_t npairs = L.pairs, NB = L.nbuf;
	const int rmask = (int)L.ring - 1;
	uint4* tab = reinterpret_cast<uint4*>(smem_raw);
	for (uint32_t i = threadIdx.x; i < 8 * 256; i += blockDim.x)
		tab[i] = g_tab[i];
	// per pair: [plane ring + zero slot][masks x NB][hit queue x2][desc x NB][full mbarriers x NB][empty mbarriers x NB]
	const uint32_t plane_bytes = (L.ring + 1u) * 256u; // [position mod ring][lane] uint2; slot `ring` = zeros
	const uint32_t pair_bytes = plane_bytes + NB * kMaskBytes + kHitWarps * kQueueCap * 4 + NB * 32;
	unsigned char* pbase = smem_raw + kTabBytes + pair * pair_bytes;
	uint2* planes = reinterpret_cast<uint2*>(pbase);
	uint32_t* hwbuf = reinterpret_cast<uint32_t*>(pbase + plane_bytes);          // [NB][16][32]
	uint32_t* queues = hwbuf + NB * (kMaskBytes / 4);                             // [kHitWarps][kQueueCap]
	BodyDesc* desc = reinterpret_cast<BodyDesc*>(queues + kHitWarps * kQueueCap);  // [NB]
	uint64_t* bar_full = reinterpret_cast<uint64_t*>(desc + NB);                   // [NB]
	uint64_t* bar_empty = bar_full + NB;                                           // [NB]
	if (role == 0 && pair < npairs) {
		planes[L.ring * 32 + lane] = make_uint2(0u, 0u);
		if (lane < NB) {
			mbar_init(&bar_full[lane], 1);
			mbar_init(&bar_empty[lane], 1);
		}
	}
	__syncthreads();

	WarpCtx c;
	c.words = words;
	c.stride = stride;
	c.k = L.k;
	c.rBits = L.rBits;
	c.tab = tab;
	c.rot_a = L.rot_a;
	c.rot_b = L.rot_b;
	c.ctr_k = ctr_k;
	c.nored = L.dbg & 2u;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(c.red_policy));
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(c.keep_policy));

	if (role != 0) {
		// ================= hit warps: consume the masks of every other hand-off unit =================
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kHitRegs));
		if (pair >= npairs)
			return;
		const uint32_t hw_id = role - 1;
		for (uint32_t it = hw_id;; it += kHitWarps) {
			const uint32_t b = it % NB;
			mbar_wait_sleep(&bar_full[b], (it / NB) & 1u);
			const BodyDesc d = desc[b];
			if (d.nq == 0)
				break;
			c.rb = d.rb;
			c.nwords = d.nwords;
			if (!(L.dbg & 1u))
				drain_body<S>(c, hwbuf + b * (kMaskBytes / 4), queues + hw_id * kQueueCap, d.nq, d.q0, lane);
			if (lane == 0)
				mbar_arrive(&bar_empty[b]);
		}
		return;
	}

	// ================= scan warps =================
	asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kScanRegs));
	if (pair >= npairs)
		return;
	const int k = (int)L.k;
	unsigned long long f1_local = 0;
	uint32_t it = 0; // hand-offs so far; unit `it` uses mask buffer it % NB and hit warp it & 1
	const uint32_t n_tiles = (n_rec + kTileRecs - 1) / kTileRecs;
	for (uint32_t tile = blockIdx.x * npairs + pair; tile < n_tiles; tile += gridDim.x * npairs) {
		const uint32_t rb = tile * kTileRecs;
		// ---- first 16 bytes of every record (length + 3 base words); is the tile uniform? ----
		// The last tile may be partial: slots past the end re-read the batch's last record (so every load is in
		// bounds and the uniformity test is unaffected) and are masked out of every hand-off (vmask).
		const uint32_t nvalid = min(kTileRecs, n_rec - rb), last_rec = n_rec - 1u;
		uint32_t vmask = 0;
		bool uniform;
		uint4 v[32];
		uint32_t len0 = 0;
		{
#pragma unroll
			for (int s = 0; s < 32; s++) {
				v[s] = ldg_keep_v4(reinterpret_cast<const uint4*>(words + (uint64_t)min(rb + s * 32u + lane, last_rec) * stride), c.keep_policy);
				vmask |= (s * 32u + lane < nvalid ? 1u : 0u) << s;
			}
			len0 = __shfl_sync(0xFFFFFFFFu, v[0].x, 0);
			bool same = true;
#pragma unroll
			for (int s = 0; s < 32; s++)
				same = same && (v[s].x == len0);
			uniform = __all_sync(0xFFFFFFFFu, same);
		}
		if (!uniform) {
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
		const uint32_t nwords = (uint32_t)(n + 15) >> 4;
		const uint32_t ngroups = (nwords + 1 + 3) / 4; // uint4 groups per record incl. the length word
		State st;
#pragma unroll
		for (int j = 0; j < 31; j++) {
			st.F[j] = L.F0[j];
			st.R[j] = L.R0[j];
		}
		// One column = one packed word of every record = 16 positions: transpose it into the plane ring, scan it,
		// hand its masks over.  The next 16 bytes of every record are requested right after the last word of the
		// current 16 bytes has been taken out of v[], so the loads fly during a whole column's scan.
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
					v[s] = ldg_keep_v4(reinterpret_cast<const uint4*>(words + (uint64_t)min(rb + s * 32u + lane, last_rec) * stride) + (g + 1), c.keep_policy);
			}
			if (w == nwords / 2) {
				// half way through: warm L2 with this warp's next tile (bulk async prefetch); earlier is too early, the
				// lines would be evicted again by the sketch traffic before they are used
				const uint64_t nrb = (uint64_t)(tile + gridDim.x * npairs) * kTileRecs;
				if (lane == 0 && nrb + kTileRecs <= n_rec) {
					const uint32_t bytes = kTileRecs * stride * 4u;
					asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(words + nrb * stride), "r"(bytes) : "memory");
				}
			}
			transpose32(A);
			const int q0 = (int)(16u * w);
			{
				uint2* dst = planes + (q0 & rmask) * 32 + lane; // a column never wraps: the ring size is a multiple of 16
#pragma unroll
				for (int j = 0; j < 16; j++)
					dst[j * 32] = make_uint2(A[2 * j], A[2 * j + 1]);
			}
			__syncwarp();
			const int nq = min(kBodyPos, n - q0);
			const bool has_windows = q0 + nq >= k; // some position of this column ends a full window
			const uint32_t b = it % NB;
			uint32_t* hw = hwbuf + b * (kMaskBytes / 4);
			if (it >= NB) // the hit warp must have drained the previous use of this buffer (columns without windows write it too)
				mbar_wait(&bar_empty[b], ((it / NB) - 1) & 1u);
#pragma unroll 1
			for (int qb = q0; qb < q0 + nq; qb += kBlock) {
				DevBlock<KM, S, 0>::run(st, planes + lane, hw + (qb - q0) * 32 + lane, qb, k, rmask, vmask);
				rotate_home(st);
			}
			if (has_windows) {
				__syncwarp();
				if (lane == 0) {
					desc[b] = BodyDesc{ rb, (uint32_t)q0, (uint32_t)nq, nwords };
					mbar_arrive(&bar_full[b]);
				}
				it++;
			}
		}
		if (lane == 0)
			f1_local += (unsigned long long)nvalid * (unsigned long long)(n - k + 1);
	}
	// tell the hit warps to stop: the next kHitWarps units (one per hit warp) carry nq = 0
#pragma unroll 1
	for (uint32_t e = 0; e < kHitWarps; e++, it++) {
		const uint32_t b = it % NB;
		if (it >= NB)
			mbar_wait(&bar_empty[b], ((it / NB) - 1) & 1u);
		__syncwarp();
		if (lane == 0) {
			desc[b] = BodyDesc{ 0, 0, 0, 0 };
			mbar_arrive(&bar_full[b]);
		}
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
	kern<<<a.grid, kThreads, a.smem_bytes, a.stream>>>(a.words, a.stride, a.n_rec, a.L, a.d_tab, a.d_params, a.ctr_k, a.f1_k);
	return cudaGetLastError();
}

} // namespace bs
} // namespace ntc
