// ntc_api.cu -- the C-ABI (include/ntcard_b200.h) over the sm_100a kernels: context, pinned
// double-buffered H2D pipeline, batch submission, finish.  No CPU fallback anywhere: without a
// usable CUDA device every compute entry point fails with NTC_ENODEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/ntcard_b200.h"
#include "bitslice_core.cuh"
#include "hit_hash.cuh"
#include "internal.h"
#include "launch.h"
#include "pipeline.h"
#include "sketch_common.cuh"

namespace {

using ntc::set_err;

#define CK(call)                                                                                         \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                         \
		if (e_ != cudaSuccess)                                                                           \
			return set_err(NTC_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

// HT(c, "label", statement): the statement, timed on the host when the context was created with NTC_HOST_TIMING=1
#define HT(c, what, stmt)                                                                                  \
	do {                                                                                                   \
		if (!(c)->host_timing) {                                                                           \
			stmt;                                                                                          \
		} else {                                                                                           \
			const auto ht0_ = std::chrono::steady_clock::now();                                            \
			stmt;                                                                                          \
			(c)->host_span(what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ht0_).count()); \
		}                                                                                                  \
	} while (0)

constexpr int NBUF = 3; // device/pinned staging ring depth

struct Stage {
	uint32_t* d_words = nullptr;
	size_t cap_words = 0;
	uint32_t* d_off = nullptr;
	size_t cap_off = 0;
	uint32_t* h_words = nullptr; // pinned staging for pageable sources
	size_t cap_h_words = 0;
	uint32_t* h_off = nullptr;
	size_t cap_h_off = 0;
	cudaEvent_t copied = nullptr;   // H2D of this slot finished (host source reusable)
	cudaEvent_t consumed = nullptr; // kernels reading this slot finished (slot reusable)
	uint64_t ticket = 0;
};

} // namespace

struct ntc_ctx {
	// Every entry point locks the context: ntRead is called concurrently on one shared sketch (one OpenMP thread per file,
	// ntcard.cpp:445), so ntc_submit must be callable from several threads.  Recursive because entry points call each other.
	std::recursive_mutex mu;
	int device = 0;
	int n_sm = 148;
	cudaStream_t stream = nullptr; // compute stream (own or caller's)
	bool own_stream = false;
	cudaStream_t copy_stream = nullptr;
	unsigned nK = 0, rBits = 0, sBits = 0, kmin = 0, kmax = 0;
	unsigned k[NTC_MAX_K] = {};
	int kernel = NTC_KERNEL_AUTO;
	unsigned gap = 0;  // -g: spaced seed with `gap` don't-care bases in the middle (general kernel only)
	// nthll mode (ntc_hll_create): HyperLogLog registers instead of the ntCard sketch; no counters, no hit log
	unsigned hll_bits = 0;
	uint8_t* d_hll = nullptr; // 2^hll_bits one-byte registers
	bool own_hll = false;
	size_t hll_bytes = 0;
	// nthll fast path: the smallest register as last seen by the host (registers only rise, so a stale value is a safe bound)
	uint32_t hll_min = 0;
	uint64_t hll_seen = 0;               // records since the last reset (sizes the chunks of the pre-filter path)
	uint32_t* d_hll_scratch = nullptr;   // [0] smallest register (hll_min_kernel), [2..3] candidate count, [4..] scan-kernel control words
	uint32_t* h_hll_min = nullptr;       // pinned
	cudaEvent_t hll_min_ev = nullptr;
	bool hll_min_pending = false;
	uint32_t mask_prefetch = 1;          // hit kernel prefetches its next unit's mask rows to L2 (NTC_MASK_PREFETCH=0: off)
	bool hll_fast = true;                // NTC_HLL_FAST=0: the 64-bit recurrence only
	uint32_t hll_grow = 200;             // next pre-filter chunk = records seen so far x hll_grow / 100 (NTC_HLL_GROW)
	uint32_t hll_first = 256u << 10;     // records per chunk while the registers are below 4, and the smallest pre-filter chunk (NTC_HLL_FIRST_K, in K records)
	uint32_t* d_counters = nullptr;
	bool own_counters = false;
	size_t n_counters = 0;
	unsigned long long* d_f1 = nullptr;
	ntc::DevParams* d_params = nullptr;
	uint4* d_bs_tab = nullptr;            // hit-path byte tables of the bit-sliced kernel
	struct KInit {                          // per-k constants of the scan / hit kernels
		uint32_t F0[31], R0[31];            // initial bit-sliced state (bitslice_core.cuh init_state)
		bool polyA_sampled;                 // ntComp samples the all-A k-mer: zero padding must not be scanned (scan_kernel.cuh, mixed tiles)
	} kinit[NTC_MAX_K];
	bool totals_overridden = false;
	uint64_t totals[NTC_MAX_K] = {};
	Stage stage[NBUF];
	uint64_t next_ticket = 1;
	// piece tables (general kernel)
	uint32_t* d_piece_first = nullptr;
	size_t cap_piece_first = 0;
	uint32_t* d_piece_rec = nullptr;
	size_t cap_piece_rec = 0;
	void* d_scan_tmp = nullptr;
	size_t cap_scan_tmp = 0;
	// re-tiling of long ragged records (sketch_kernels.cu)
	uint32_t* d_rt_counts = nullptr;   // [3][n_rec + 1]: full pieces, tail flags, tail words (prefix sums)
	size_t cap_rt_counts = 0;
	uint32_t* d_rt_uniform = nullptr;
	size_t cap_rt_uniform = 0;
	uint32_t* d_rt_tail_words = nullptr;
	size_t cap_rt_tail_words = 0;
	uint32_t* d_rt_tail_off = nullptr;
	size_t cap_rt_tail_off = 0;
	uint32_t* d_rt_off = nullptr;      // offsets made up for uniform-stride batches of long records
	size_t cap_rt_off = 0;
	uint64_t n_retiled = 0;
	uint64_t n_padded = 0;             // ragged batches of short records padded to a uniform stride for the pipeline
	// sketch pipeline (scan -> hit log -> apply), pipeline.h
	ntc::pl::Pool pool{};
	uint32_t* d_pool_ctl_region = nullptr; // ctl + slice_nblk + zero_done + apply_done + cand (zeroed by reset); the head of d_log_slab
	size_t pool_ctl_bytes = 0;
	uint32_t* d_log_slab = nullptr;        // one allocation (one IPC handle): [control region, padded to 4 KB][slice_blocks]
	size_t log_slab_ctl_words = 0;         // words of the padded control region = offset of slice_blocks in the slab
	// multi-GPU reduction over peer memory (ntc_peer_attach): the logs of all ranks of the node
	ntc::pl::PeerLogs peers{};
	int peer_world = 0, peer_rank = 0;
	void* peer_mapped[2 * ntc::pl::kMaxPeers] = {}; // cudaIpcOpenMemHandle results (to close)
	uint32_t* d_masks = nullptr;
	size_t cap_masks = 0;
	uint32_t* d_tile_info = nullptr;
	size_t cap_tile_info = 0;
	uint32_t* d_offchk = nullptr; // ntc_submit_device: {longest record, bad offsets} of a ragged batch
	bool pending = false;     // the hit log may hold entries, or the sketch is not materialised yet: flush before reading it
	bool use_pipeline = true;
	bool use_fused = false;    // NTC_FUSED=1: the fused sketch kernel (fused_kernel.cuh) instead of scan + hit; exact, but measured slower
	unsigned fused_dbg = 0;    // NTC_FUSED_DBG: timing experiments (fused_kernel.cuh), results are wrong when set
	bool no_stage = false;     // NTC_NO_STAGE=1 (stand-alone hit kernel: gather with __ldg instead of a TMA-staged tile)
	bool no_retile = false;    // NTC_NO_RETILE=1
	bool clear_by_memset = true;
	bool pad_ragged = true;    // NTC_PAD=0: NTC_KERNEL_AUTO leaves short ragged batches to the general kernel instead of padding them for the pipeline
	// NTC_HOST_TIMING=1: host time spent in the calls of the submit path, per call site, printed by ntc_destroy (diagnosis only)
	bool host_timing = false;
	struct HostSpan { const char* what; double total_ms, max_ms; uint64_t n; };
	std::vector<HostSpan> host_spans;
	void host_span(const char* what, double ms)
	{
		for (auto& h : host_spans)
			if (h.what == what) {
				h.total_ms += ms;
				h.max_ms = std::max(h.max_ms, ms);
				h.n++;
				return;
			}
		host_spans.push_back({ what, ms, ms, 1 });
	}
	unsigned scan_prefetch = 2; // scan kernel: L2 prefetch of a warp's next tile one column before the end of the current one (NTC_SCAN_PREFETCH: 0 = none, 1 = half way: measured, thrashes L2)
	bool partial = false;     // after ntc_flush_slices: only the owned slices of the sketch are defined (until ntc_reset)
	std::vector<uint32_t> h_nblk; // block counts per slice as of the last ntc_log_counts
	uint32_t log_used = 0;
	static constexpr int kRunSlots = 8;
	struct RunSlot {
		uint32_t *h = nullptr, *d = nullptr;
		size_t cap_h = 0, cap_d = 0;
		cudaEvent_t done = nullptr;
	} run_slot[kRunSlots];
	unsigned next_run_slot = 0;
	unsigned apply_grid = 0, hit_grid_max = 0;
	unsigned chunk_waves = 0; // NTC_CHUNK_WAVES: scan waves per pipeline chunk (0 = as large as the hit-log pool allows)
	unsigned multik_chunk_mb = 0; // NTC_MULTIK_CHUNK_MB: with several k, cap the chunks at this many MB of packed reads (L2 reuse across k)
	uint64_t n_flush_launches = 0;
	// finish buffers
	uint16_t* d_narrow = nullptr;
	uint32_t* d_phist = nullptr;
	// stats / timing
	uint64_t n_launches = 0, n_batches = 0;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing; // per batch, around the sketch kernels
	struct StageSpan { cudaEvent_t a, b; int stage; };
	std::vector<StageSpan> stage_timing;                      // per pipeline stage: 0 scan, 1 hit (+ fallback), 2 apply
	double stage_ms[3] = { 0, 0, 0 };
	std::vector<cudaEvent_t> event_pool;
	double kernel_ms = 0;
	uint64_t n_timed = 0;
};

namespace {

int use_device(ntc_ctx* c)
{
	CK(cudaSetDevice(c->device));
	return NTC_OK;
}

template <typename T>
int grow(T** p, size_t* cap, size_t need, bool pinned_host)
{
	if (need <= *cap)
		return NTC_OK;
	size_t n = std::max(need, *cap + *cap / 2);
	if (*p) {
		if (pinned_host)
			CK(cudaFreeHost(*p));
		else
			CK(cudaFree(*p));
		*p = nullptr;
		*cap = 0;
	}
	if (pinned_host)
		CK(cudaHostAlloc((void**)p, n * sizeof(T), cudaHostAllocDefault));
	else
		CK(cudaMalloc((void**)p, n * sizeof(T)));
	*cap = n;
	return NTC_OK;
}

int get_event(ntc_ctx* c, cudaEvent_t* ev)
{
	if (!c->event_pool.empty()) {
		*ev = c->event_pool.back();
		c->event_pool.pop_back();
		return NTC_OK;
	}
	CK(cudaEventCreate(ev));
	return NTC_OK;
}

void build_params(const ntc_ctx* c, ntc::DevParams* P)
{
	memset(P, 0, sizeof *P);
	P->nK = c->nK;
	P->rBits = c->rBits;
	P->sBits = c->sBits;
	P->kmin = c->kmin;
	for (unsigned ki = 0; ki < c->nK; ki++) {
		const unsigned k = c->k[ki];
		P->k[ki] = k;
		for (unsigned in = 0; in < 4; in++)
			for (unsigned out = 0; out < 4; out++) {
				// NTF64 / NTR64 sliding forms, nthash.hpp:242-257
				P->tab[ki].xf[in | out << 2] = ntc::seed_of(in) ^ ntc::srol_n(ntc::seed_of(out), k);
				P->tab[ki].xr[in | out << 2] = ntc::srol_n(ntc::seed_of(3 - in), k) ^ ntc::seed_of(3 - out);
			}
	}
	P->gap = c->gap;
	if (c->gap) { // gap seeds: rolling constants of the inner window, rotation a = (k - gap) / 2 (sketch_common.cuh)
		const unsigned a = (c->k[0] - c->gap) / 2, g = c->gap;
		P->gap_a31 = a % 31;
		P->gap_a33 = a % 33;
		for (unsigned in = 0; in < 4; in++)
			for (unsigned out = 0; out < 4; out++) {
				P->gtab.xf[in | out << 2] = ntc::seed_of(in) ^ ntc::srol_n(ntc::seed_of(out), g);
				P->gtab.xr[in | out << 2] = ntc::srol_n(ntc::seed_of(3 - in), g) ^ ntc::seed_of(3 - out);
			}
	}
}

// ---- sketch pipeline -------------------------------------------------------------------------------------
int pool_create(ntc_ctx* c)
{
	ntc::pl::Pool& P = c->pool;
	const char* env = getenv("NTC_POOL_BLOCKS");
	// default: 768 Ki blocks (x 256 entries x 4 B = 768 MiB) per k -- a flush walks the whole sketch (1 GiB per k), so the more
	// hits fit between two flushes the better, and HBM is plentiful (100 M reads x 4 k: 48.5 ms per pass with 512 MiB, 33.6 ms
	// with 4 GiB); capped so that the block lists (n_slices x n_blocks words) stay below 4 GiB
	P.n_blocks = env ? (uint32_t)strtoul(env, nullptr, 10) : (768u << 10) * std::min(c->nK, 8u);
	if (!env) // no more log entries than twice the counters (small sketches in tests: a small pool)
		P.n_blocks = (uint32_t)std::min<uint64_t>(P.n_blocks, std::max<uint64_t>(4096, (((uint64_t)c->nK * NTC_NSAMP) << c->rBits) / 128));
	P.n_blocks = std::min(std::max(P.n_blocks, 64u), (1u << 23) - 1u); // a block id has 23 bits in the slice lists
	P.rBits = c->rBits;
	P.nK = c->nK;
	P.ahead = getenv("NTC_APPLY_AHEAD") ? (uint32_t)atoi(getenv("NTC_APPLY_AHEAD")) : 2u;
	const uint32_t idx_bits = c->rBits + 1;
	{
		// 2^23 counters = 32 MiB per slice, 32 slices per k at r = 27.  Measured (profiles/r02_apply_tuning.txt): on one GPU 16 MiB slices
		// are 2 % faster per step (apply 0.347 vs 0.371 ms), but the multi-GPU reduction pays a fixed cost per owned slice (remote round
		// trips, hand-offs): 8 GPUs 7.9e12 k-mers/s with 32 MiB slices against 7.1e12 with 16 MiB; 64 MiB: apply 0.74 ms (L2 footprint)
		const uint32_t ss = getenv("NTC_SLICE_SHIFT") ? (uint32_t)atoi(getenv("NTC_SLICE_SHIFT")) : 23u;
		P.bin_shift = idx_bits <= ss ? idx_bits : std::max(ss, idx_bits - 6u); // <= 64 slices per k
	}
	P.nbins = 1u << (idx_bits - P.bin_shift);
	P.n_slices = c->nK * P.nbins;
	if (!env)
		P.n_blocks = std::max(64u, std::min(P.n_blocks, (1u << 30) / P.n_slices));
	// every slice's block list can hold the whole pool: a list can then never fill up before the pool does, so the only
	// "out of space" event is pool exhaustion, which both hit paths handle exactly (deferred tiles / conditional flush)
	P.slice_cap = P.n_blocks;
	CK(cudaMalloc((void**)&P.entries, (size_t)P.n_blocks * ntc::pl::kBlkEntries * sizeof(uint32_t)));
	// control region: [ctl CTL_WORDS][slice_nblk n_slices][zero_done n_slices][apply_done n_slices][pad][cand 2 words], padded to
	// 4 KB, then the block lists -- ONE allocation, so that one IPC handle maps everything a peer needs besides the pool
	const size_t words = ntc::pl::CTL_WORDS + 3 * (size_t)P.n_slices + 4;
	c->pool_ctl_bytes = words * sizeof(uint32_t);
	c->log_slab_ctl_words = (words + 1023) & ~(size_t)1023;
	CK(cudaMalloc((void**)&c->d_log_slab, (c->log_slab_ctl_words + (size_t)P.n_slices * P.slice_cap) * sizeof(uint32_t)));
	c->d_pool_ctl_region = c->d_log_slab;
	P.slice_blocks = c->d_log_slab + c->log_slab_ctl_words;
	P.ctl = c->d_pool_ctl_region;
	P.slice_nblk = P.ctl + ntc::pl::CTL_WORDS;
	P.zero_done = P.slice_nblk + P.n_slices;
	P.apply_done = P.zero_done + P.n_slices;
	P.cand = reinterpret_cast<unsigned long long*>(P.apply_done + P.n_slices + ((ntc::pl::CTL_WORDS + 3 * P.n_slices) & 1u));
	c->apply_grid = (unsigned)ntc::pl::apply_max_grid(c->n_sm);
	c->hit_grid_max = (unsigned)c->n_sm * 2u;
	P.max_groups = (unsigned)c->n_sm * 8u; // fused kernel: one saved state per warp (8 per SM); hit kernel: per group (<= 4 per SM)
	P.epoch = 1;
	{
		const size_t gwords = (size_t)c->nK * P.max_groups * ntc::pl::gstate_row(P.nbins);
		CK(cudaMalloc((void**)&P.gstate, gwords * sizeof(uint32_t)));
		CK(cudaMemset(P.gstate, 0, gwords * sizeof(uint32_t)));
	}
	return NTC_OK;
}

// CUDA events around one pipeline stage (0 scan, 1 hit + fallback, 2 apply) on the compute stream
int stage_begin(ntc_ctx* c, int stage)
{
	ntc_ctx::StageSpan sp;
	int rc;
	if ((rc = get_event(c, &sp.a)) || (rc = get_event(c, &sp.b)))
		return rc;
	sp.stage = stage;
	CK(cudaEventRecord(sp.a, c->stream));
	c->stage_timing.push_back(sp);
	return NTC_OK;
}
int stage_end(ntc_ctx* c)
{
	CK(cudaEventRecord(c->stage_timing.back().b, c->stream));
	return NTC_OK;
}

// Apply everything that is pending to the counters in HBM (and materialise them after a reset).
int flush(ntc_ctx* c)
{
	c->h_nblk.clear(); // (ntc_log_export / ntc_log_import refuse to work from a stale snapshot)
	if (!c->pending)
		return NTC_OK;
	cudaEvent_t e0, e1;
	int rc;
	if ((rc = get_event(c, &e0)) || (rc = get_event(c, &e1)))
		return rc;
	CK(cudaEventRecord(e0, c->stream));
	if ((rc = stage_begin(c, 2)))
		return rc;
	CK(ntc::pl::launch_apply(c->pool, c->d_counters, 1, 0, c->apply_grid, c->stream));
	if ((rc = stage_end(c)))
		return rc;
	CK(cudaEventRecord(e1, c->stream));
	c->timing.emplace_back(e0, e1);
	c->n_launches++;
	c->n_flush_launches++;
	c->pending = false;
	return NTC_OK;
}

struct PipeShape {
	uint32_t ring = 0, nwarps = 0, npos_max = 0, rows_per_unit = 64, units_per_tile = 1, tiles_per_unit = 1, start_limit = 0;
	size_t smem = 0;
	uint32_t qlane = 0; // fused kernel: candidate slots per lane
};

// Shared-memory shape of the fused kernel for one k: as many warps as fit (<= 8) with a candidate queue that holds a tile's
// expected candidates per lane (npos * 32 / 2^(sBits-1)) with a wide margin; the index buffer of the append phase aliases the
// plane ring, hence qlane <= 2 * (ring + 3), and the ring may be chosen larger than k + 16 to make room.
bool fused_shape(uint32_t k, uint32_t sBits, uint32_t npos_max, uint32_t nbins, PipeShape* s)
{
	if (npos_max > ntc::pl::kFusedMaxPos)
		return false;
	const uint32_t ring_min = (k + 16 + 15) & ~15u;
	const double expect = (double)npos_max * 32.0 / (double)(1u << (sBits - 1));
	const uint32_t target = std::max(32u, ((uint32_t)(expect * 1.5) + 24u + 7u) & ~7u);
	const uint32_t qmin = std::max(32u, (uint32_t)(expect * 1.2) + 16u);
	for (uint32_t nw = ntc::pl::kFusedWarps; nw >= 1; nw--) {
		uint32_t best_q = 0, best_ring = 0;
		for (uint32_t ring = ring_min; ring <= ring_min + 256; ring += 16) {
			const size_t fixed = ntc::pl::fused_smem_bytes(ring, nw, 0, nbins);
			if (fixed > ntc::pl::kSmemMax)
				break;
			const uint32_t q_mem = (uint32_t)((ntc::pl::kSmemMax - fixed) / nw / 64);
			const uint32_t q = std::min(std::min(target, 2u * (ring + 3u)), q_mem);
			if (q > best_q) {
				best_q = q;
				best_ring = ring;
			}
			if (q >= target)
				break;
		}
		if (best_q >= qmin || (nw == 1 && best_q >= 32)) {
			s->ring = best_ring;
			s->nwarps = nw;
			s->qlane = best_q;
			s->smem = ntc::pl::fused_smem_bytes(best_ring, nw, best_q, nbins);
			return s->smem <= ntc::pl::kSmemMax;
		}
	}
	return false;
}

// Which k indices the scan -> hit -> apply pipeline can take for this batch.
uint32_t pipeline_config(const ntc_ctx* c, const ntc::BatchView& b, bool record_is_piece, PipeShape* shape, uint32_t min_rec = 1024)
{
	if (!c->use_pipeline || c->gap || c->kernel == NTC_KERNEL_ROLL64 || b.off || !record_is_piece || b.stride < 4 || (b.stride & 3u) || b.n_rec < min_rec ||
	    (reinterpret_cast<uintptr_t>(b.words) & 15u))
		return 0;
	uint32_t kmask = 0;
	const uint32_t max_len = (b.stride - 1) * 16;
	for (unsigned ki = 0; ki < c->nK; ki++) {
		const unsigned k = c->k[ki];
		if (k >= 288 || !ntc::pl::have_scan_kernel(k, c->sBits))
			continue;
		PipeShape& s = shape[ki];
		if (max_len < k) { // no k-mer fits any record of this batch: nothing to do for this k
			s.npos_max = 0;
			kmask |= 1u << ki;
			continue;
		}
		s.npos_max = max_len - k + 1;
		if (s.npos_max > 65535)
			continue;
		if (c->use_fused) {
			if (!fused_shape(k, c->sBits, s.npos_max, c->pool.nbins, &s))
				continue;
			kmask |= 1u << ki;
			continue;
		}
		const uint32_t r = (k + 16 + 15) & ~15u; // plane ring: k + 16 positions, rounded up to whole columns
		const size_t per_warp = (size_t)(r + 3) * 256; // ring + 3 mirror slots (scan_kernel.cuh)
		const uint32_t nw = (uint32_t)std::min<size_t>(8, ntc::pl::kSmemMax / per_warp);
		if (nw < 1)
			continue;
		s.ring = r;
		s.nwarps = nw;
		s.smem = nw * per_warp;
		// a mask row holds 32 x 32 k-mers, sampled at 2^-(sBits-1): 2^(sBits-1) rows = 1024 candidates on average.
		// A unit of the hit kernel is either a run of rows of one tile or several whole tiles.
		const uint32_t target = std::min(4096u, std::max(4u, 2u << (c->sBits - 1))); // 2048 candidates per unit
		if (s.npos_max >= target) {
			s.units_per_tile = std::max(1u, (s.npos_max + target / 2) / target);
			s.rows_per_unit = (s.npos_max + s.units_per_tile - 1) / s.units_per_tile;
			s.tiles_per_unit = 1;
		} else {
			s.units_per_tile = 1;
			s.tiles_per_unit = std::max(1u, target / s.npos_max);
			s.rows_per_unit = s.npos_max;
		}
		kmask |= 1u << ki;
	}
	return kmask;
}

int run_pipeline_chunk(ntc_ctx* c, const ntc::BatchView& b, unsigned ki, const PipeShape& sh);

// All pipeline k (kmask) over one batch, chunk by chunk.  A chunk is a whole number of tiles and is kept small enough that the
// hits it is expected to produce for any one k (tile k-mers / 2^(sBits-1)) fill at most ~40 % of the hit-log pool: a batch whose
// hits exceed the pool would otherwise run the second half of its hit kernel on direct increments into HBM (~8x slower;
// 100 M reads x k=32 in ONE batch: 186 M hits against 134 M pool entries).  The k loop is INSIDE the chunk loop -- with
// NTC_MULTIK_CHUNK_MB=<n> the chunks are additionally capped at n MB of packed reads, so that the second and later k find them
// in L2 and the reads cross DRAM once for all k (ntRead re-walks the read per k, ntcard.cpp:150; measured: DESIGN section 8).
int run_pipeline_all(ntc_ctx* c, const ntc::BatchView& b, uint32_t kmask, const PipeShape* shape)
{
	const uint32_t n_tiles = (b.n_rec + 1023) / 1024;
	const double pool_entries = (double)c->pool.n_blocks * ntc::pl::kBlkEntries;
	uint64_t chunk_tiles = n_tiles;
	unsigned nk = 0;
	for (unsigned ki = 0; ki < c->nK; ki++)
		if (((kmask >> ki) & 1u) && shape[ki].npos_max) {
			nk++;
			const double per_tile = 1024.0 * shape[ki].npos_max / (double)(1u << (c->sBits - 1));
			chunk_tiles = std::min<uint64_t>(chunk_tiles, std::max<uint64_t>(1, (uint64_t)(0.4 * pool_entries / per_tile)));
			if (c->chunk_waves)
				chunk_tiles = std::min<uint64_t>(chunk_tiles, (uint64_t)c->n_sm * std::max(1u, shape[ki].nwarps) * c->chunk_waves);
		}
	if (nk == 0)
		return NTC_OK;
	if (c->multik_chunk_mb && nk > 1)
		chunk_tiles = std::min<uint64_t>(chunk_tiles, std::max<uint64_t>(1, ((uint64_t)c->multik_chunk_mb << 20) / (1024ull * b.stride * 4)));
	// equal chunks
	const uint64_t n_chunks = (n_tiles + chunk_tiles - 1) / chunk_tiles;
	chunk_tiles = (n_tiles + n_chunks - 1) / n_chunks;
	for (uint64_t t0 = 0; t0 < n_tiles; t0 += chunk_tiles) {
		ntc::BatchView sub = b;
		const uint64_t r0 = t0 * 1024;
		sub.words = b.words + r0 * b.stride;
		sub.n_rec = (uint32_t)std::min<uint64_t>(chunk_tiles * 1024, b.n_rec - r0);
		sub.n_words = (uint64_t)sub.n_rec * b.stride;
		for (unsigned ki = 0; ki < c->nK; ki++)
			if (((kmask >> ki) & 1u) && shape[ki].npos_max) {
				int rc = run_pipeline_chunk(c, sub, ki, shape[ki]);
				if (rc)
					return rc;
			}
	}
	return NTC_OK;
}

// One k over one batch with the fused kernel: pass 0 -> conditional flush -> pass 1 (deferred tiles) -> fallback (flagged tiles).
int run_fused_chunk(ntc_ctx* c, const ntc::BatchView& b, unsigned ki, const PipeShape& sh)
{
	int rc;
	const uint32_t n_tiles = (b.n_rec + 1023) / 1024;
	if ((rc = grow(&c->d_tile_info, &c->cap_tile_info, (size_t)n_tiles, false)))
		return rc;
	ntc::pl::Pool& P = c->pool;
	CK(cudaMemsetAsync(P.ctl + ntc::pl::CTL_NFLAG, 0, 2 * sizeof(uint32_t), c->stream)); // NFLAG and NDEFER
	ntc::pl::FusedArgs fa;
	fa.words = b.words;
	fa.stride = b.stride;
	fa.n_rec = b.n_rec;
	fa.n_tiles = n_tiles;
	fa.L.k = c->k[ki];
	fa.L.ring = sh.ring;
	fa.L.nwarps = sh.nwarps;
	fa.L.npos_max = sh.npos_max;
	fa.L.start_limit = sh.start_limit;
	fa.L.mixed_ok = c->kinit[ki].polyA_sampled ? 0u : 1u;
	fa.L.prefetch = c->scan_prefetch;
	memcpy(fa.L.F0, c->kinit[ki].F0, sizeof fa.L.F0);
	memcpy(fa.L.R0, c->kinit[ki].R0, sizeof fa.L.R0);
	fa.ki = ki;
	fa.sBits = c->sBits;
	fa.pass = 0;
	fa.qlane = sh.qlane;
	fa.dbg = c->fused_dbg;
	fa.d_tab = c->d_bs_tab;
	fa.hk = ntc::pl::make_hashk(c->k[ki]);
	fa.ctr_k = c->d_counters + ((size_t)ki * NTC_NSAMP << c->rBits);
	fa.pool = P;
	fa.tile_info = c->d_tile_info;
	fa.f1_k = c->d_f1 + ki;
	fa.grid = std::min<unsigned>((unsigned)c->n_sm, n_tiles);
	fa.smem_bytes = sh.smem;
	fa.stream = c->stream;
	if ((rc = stage_begin(c, 0)))
		return rc;
	CK(ntc::pl::launch_fused(c->k[ki], c->sBits, fa));
	if ((rc = stage_end(c)) || (rc = stage_begin(c, 1)))
		return rc;
	// Normally the next three launches find nothing to do and leave at once: the flush runs only when pass 0 deferred tiles (pool
	// exhausted before the sketch was materialised) or flagged tiles (mixed lengths whose padding k-mer is sampled), pass 1 and the
	// fallback kernel only over those tiles.
	CK(ntc::pl::launch_apply(P, c->d_counters, 2, 0, c->apply_grid, c->stream));
	fa.pass = 1;
	CK(ntc::pl::launch_fused(c->k[ki], c->sBits, fa));
	CK(ntc::pl::launch_fallback(b.words, b.stride, b.n_rec, n_tiles, c->d_tile_info, c->d_params, ki, fa.ctr_k, P.ctl, c->n_sm, c->stream));
	if ((rc = stage_end(c)))
		return rc;
	c->n_launches += 4;
	c->pending = true;
	return NTC_OK;
}

int run_pipeline_chunk(ntc_ctx* c, const ntc::BatchView& b, unsigned ki, const PipeShape& sh)
{
	if (c->use_fused)
		return run_fused_chunk(c, b, ki, sh);
	int rc;
	const uint32_t n_tiles = (b.n_rec + 1023) / 1024;
	HT(c, "chunk: grow masks / tile_info", rc = grow(&c->d_masks, &c->cap_masks, (size_t)n_tiles * sh.npos_max * 32, false);
	   if (!rc) rc = grow(&c->d_tile_info, &c->cap_tile_info, (size_t)n_tiles, false));
	if (rc)
		return rc;
	ntc::pl::Pool& P = c->pool;
	if (c->clear_by_memset) {
		HT(c, "chunk: 2 x cudaMemsetAsync", CK(cudaMemsetAsync(P.cand, 0, sizeof(unsigned long long), c->stream));
		   CK(cudaMemsetAsync(P.ctl + ntc::pl::CTL_NFLAG, 0, sizeof(uint32_t), c->stream)));
	} else {
		HT(c, "chunk: launch_batch_clear", CK(ntc::pl::launch_batch_clear(P, c->stream)));
	}
	ntc::pl::ScanArgs sa;
	sa.words = b.words;
	sa.stride = b.stride;
	sa.n_rec = b.n_rec;
	sa.L.k = c->k[ki];
	sa.L.ring = sh.ring;
	sa.L.nwarps = sh.nwarps;
	sa.L.npos_max = sh.npos_max;
	sa.L.start_limit = sh.start_limit;
	sa.L.mixed_ok = c->kinit[ki].polyA_sampled ? 0u : 1u;
	sa.L.prefetch = c->scan_prefetch;
	memcpy(sa.L.F0, c->kinit[ki].F0, sizeof sa.L.F0);
	memcpy(sa.L.R0, c->kinit[ki].R0, sizeof sa.L.R0);
	sa.masks = c->d_masks;
	sa.tile_info = c->d_tile_info;
	sa.f1_k = c->d_f1 + ki;
	sa.cand = P.cand;
	sa.ctl = P.ctl;
	sa.grid = std::min<unsigned>((unsigned)c->n_sm, n_tiles);
	sa.smem_bytes = sh.smem;
	sa.stream = c->stream;
	if ((rc = stage_begin(c, 0)))
		return rc;
	HT(c, "chunk: launch_scan", CK(ntc::pl::launch_scan(c->k[ki], c->sBits, sa)));
	if ((rc = stage_end(c)) || (rc = stage_begin(c, 1)))
		return rc;
	ntc::pl::HitArgs ha;
	ha.words = b.words;
	ha.stride = b.stride;
	ha.n_rec = b.n_rec;
	ha.n_tiles = n_tiles;
	ha.k = c->k[ki];
	ha.ki = ki;
	ha.sBits = c->sBits;
	ha.npos_max = sh.npos_max;
	ha.mask_prefetch = c->mask_prefetch;
	ha.rows_per_unit = sh.rows_per_unit;
	ha.units_per_tile = sh.units_per_tile;
	ha.tiles_per_unit = sh.tiles_per_unit;
	ha.n_units = sh.tiles_per_unit > 1 ? (n_tiles + sh.tiles_per_unit - 1) / sh.tiles_per_unit : n_tiles * sh.units_per_tile;
	ha.masks = c->d_masks;
	ha.tile_info = c->d_tile_info;
	ha.d_tab = c->d_bs_tab;
	ha.hk = ntc::pl::make_hashk(c->k[ki]);
	ha.ctr_k = c->d_counters + ((size_t)ki * NTC_NSAMP << c->rBits);
	ha.pool = P;
	ha.stream = c->stream;
	const bool staged = ntc::pl::hit_can_stage(b.stride, sh.units_per_tile, sh.tiles_per_unit) && !c->no_stage;
	const unsigned gpc = staged ? 3u : 2u; // groups per CTA
	const unsigned hit_ctas = std::min<unsigned>((unsigned)c->n_sm * (staged ? 1u : 2u), (ha.n_units + gpc - 1) / gpc);
	// conditional flush: only when this batch could exhaust the pool while the sketch is not materialised yet
	HT(c, "chunk: launch_apply (conditional)", CK(ntc::pl::launch_apply(P, c->d_counters, 0, 3 * P.max_groups * P.nbins + 8, c->apply_grid, c->stream)));
	HT(c, "chunk: launch_hit", CK(ntc::pl::launch_hit(ha, staged, hit_ctas)));
	HT(c, "chunk: launch_fallback", CK(ntc::pl::launch_fallback(b.words, b.stride, b.n_rec, n_tiles, c->d_tile_info, c->d_params, ki, ha.ctr_k, P.ctl, c->n_sm, c->stream)));
	if ((rc = stage_end(c)))
		return rc;
	c->n_launches += 4;
	c->pending = true;
	return NTC_OK;
}

// Piece tables of a batch whose records may be longer than one piece (general kernel, HLL kernel).
int prepare_pieces(ntc_ctx* c, const ntc::BatchView& b, uint64_t* bound_out)
{
	int rc;
	const uint64_t bound = (uint64_t)b.n_rec + (16 * b.n_words) / ntc::PIECE_STARTS + 1;
	if (bound > 0xFFFFFFFFull)
		return set_err(NTC_EINVAL, "batch too large: %llu pieces", (unsigned long long)bound);
	if ((rc = grow(&c->d_piece_first, &c->cap_piece_first, (size_t)b.n_rec + 1, false)) ||
	    (rc = grow(&c->d_piece_rec, &c->cap_piece_rec, (size_t)bound, false)))
		return rc;
	size_t tmp = ntc::piece_scan_temp_bytes(b.n_rec + 1);
	char* t = (char*)c->d_scan_tmp;
	if ((rc = grow(&t, &c->cap_scan_tmp, tmp, false)))
		return rc;
	c->d_scan_tmp = t;
	CK(ntc::launch_piece_tables(b, c->kmin, c->d_piece_first, c->d_piece_rec, c->d_scan_tmp, c->cap_scan_tmp, c->stream));
	c->n_launches += 3;
	*bound_out = bound;
	return NTC_OK;
}

// The general kernel over one batch for the k indices in kmask (direct increments: the caller has flushed).
int run_roll64(ntc_ctx* c, const ntc::BatchView& b, bool record_is_piece, uint32_t kmask)
{
	int rc;
	uint64_t bound = 0;
	if (!record_is_piece && (rc = prepare_pieces(c, b, &bound)))
		return rc;
	HT(c, "batch: launch_roll64", CK(ntc::launch_roll64(b, record_is_piece, bound, c->d_piece_first, c->d_piece_rec, c->d_params, c->d_counters,
	    c->d_f1, kmask, c->n_sm, c->stream)));
	c->n_launches += 1;
	return NTC_OK;
}

// nthll mode: every canonical hash of the batch into the HyperLogLog registers (hll_kernels.cu)
int run_hll_general(ntc_ctx* c, const ntc::BatchView& b, bool record_is_piece)
{
	int rc;
	uint64_t bound = 0;
	if (!record_is_piece && (rc = prepare_pieces(c, b, &bound)))
		return rc;
	CK(ntc::launch_hll(b, record_is_piece, bound, c->d_piece_first, c->d_piece_rec, c->d_params, c->hll_bits, c->d_hll, c->d_f1, c->n_sm,
	    c->stream));
	c->n_launches += 1;
	return NTC_OK;
}

// smallest register -> pinned host word, asynchronously; hll_min_poll() picks it up when it has arrived
int hll_min_request(ntc_ctx* c)
{
	if (!c->d_hll_scratch) {
		CK(cudaMalloc((void**)&c->d_hll_scratch, 16 * sizeof(uint32_t)));
		CK(cudaMemset(c->d_hll_scratch, 0, 16 * sizeof(uint32_t)));
		CK(cudaHostAlloc((void**)&c->h_hll_min, sizeof(uint32_t), cudaHostAllocDefault));
		CK(cudaEventCreateWithFlags(&c->hll_min_ev, cudaEventDisableTiming));
	}
	CK(ntc::launch_hll_min(c->d_hll, c->hll_bits, c->d_hll_scratch, c->stream));
	CK(cudaMemcpyAsync(c->h_hll_min, c->d_hll_scratch, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaEventRecord(c->hll_min_ev, c->stream));
	c->hll_min_pending = true;
	c->n_launches++;
	return NTC_OK;
}
int hll_min_poll(ntc_ctx* c, bool wait)
{
	if (!c->hll_min_pending)
		return NTC_OK;
	cudaError_t e = wait ? cudaEventSynchronize(c->hll_min_ev) : cudaEventQuery(c->hll_min_ev);
	if (e == cudaErrorNotReady) {
		cudaGetLastError();
		return NTC_OK;
	}
	CK(e);
	c->hll_min = std::max(c->hll_min, *c->h_hll_min);
	c->hll_min_pending = false;
	return NTC_OK;
}

// nthll mode: every canonical hash of the batch into the HyperLogLog registers (hll_kernels.cu).  Uniform-stride batches take
// the bit-sliced pre-filter (scan kernel, "top T bits zero") + hll_hit_kernel as soon as every register has reached 4;
// until then -- the first 256 K reads of a run at 2^16 registers -- chunks go through the 64-bit recurrence.
int run_hll(ntc_ctx* c, const ntc::BatchView& b, bool record_is_piece)
{
	int rc;
	const unsigned k = c->k[0];
	const uint32_t max_len = b.off ? 0 : (b.stride - 1) * 16;
	const bool eligible = c->hll_fast && c->use_pipeline && !b.off && record_is_piece && b.stride >= 4 && !(b.stride & 3u) && b.n_rec >= 4096 &&
	                      !(reinterpret_cast<uintptr_t>(b.words) & 15u) && k < 288 && max_len >= k && max_len - k + 1 <= 65535 &&
	                      ntc::pl::have_scan_kernel(k, 7);
	if (!eligible)
		return run_hll_general(c, b, record_is_piece);
	const uint32_t ring = (k + 16 + 15) & ~15u;
	const size_t per_warp = (size_t)(ring + 3) * 256;
	const uint32_t nw = (uint32_t)std::min<size_t>(8, ntc::pl::kSmemMax / per_warp);
	if (nw < 1)
		return run_hll_general(c, b, record_is_piece);
	const uint32_t npos_max = max_len - k + 1;
	if (!c->own_hll) {
		// caller-owned registers (a torch tensor): the caller may have cleared or replaced them between two batches, so what the
		// host and the device word remember of their minimum is not to be trusted -- one look per batch (a stream sync)
		if ((rc = hll_min_poll(c, true)))
			return rc;
		c->hll_min = 0;
		if ((rc = hll_min_request(c)) || (rc = hll_min_poll(c, true)))
			return rc;
	}
	uint32_t done = 0;
	while (done < b.n_rec) {
		if ((rc = hll_min_poll(c, false)))
			return rc;
		ntc::BatchView sub = b;
		sub.words = b.words + (uint64_t)done * b.stride;
		if (c->hll_min < 4) {
			// registers still low: 256 K records through the 64-bit recurrence, then look at the smallest register again
			sub.n_rec = std::min<uint32_t>(b.n_rec - done, c->hll_first);
			sub.n_words = (uint64_t)sub.n_rec * b.stride;
			if ((rc = run_hll_general(c, sub, true)) || (rc = hll_min_request(c)) || (rc = hll_min_poll(c, true)))
				return rc;
			done += sub.n_rec;
			c->hll_seen += sub.n_rec;
			continue;
		}
		// T = 13 once the host knows every register is >= 12; until then the scan kernel reads the smallest register from the
		// device (written by hll_min_kernel after the previous chunk, no host round trip) and filters at 5, 7, 9, 11 or 13 bits.
		// A chunk is twice the records seen so far (the smallest register gains two bits per 4x records; measured on 10 M reads,
		// k = 32: factor 0.5 -> 1.35 ms, 1 -> 1.21, 2 -> 1.11, 4 -> 1.06, 8 -> 1.04, 16 -> 1.08; k = 64 prefers finer chunks).
		const unsigned T = c->hll_min >= 12 ? 13u : 0u;
		sub.n_rec = b.n_rec - done;
		if (!T)
			sub.n_rec = (uint32_t)std::min<uint64_t>(sub.n_rec, std::max<uint64_t>(c->hll_first, c->hll_seen * c->hll_grow / 100));
		if (b.n_rec - done - sub.n_rec < 4096)
			sub.n_rec = b.n_rec - done; // no crumbs
		sub.n_words = (uint64_t)sub.n_rec * b.stride;
		const uint32_t n_tiles = (sub.n_rec + 1023) / 1024;
		if ((rc = grow(&c->d_masks, &c->cap_masks, (size_t)n_tiles * npos_max * 32, false)) ||
		    (rc = grow(&c->d_tile_info, &c->cap_tile_info, (size_t)n_tiles, false)))
			return rc;
		ntc::pl::ScanArgs sa;
		sa.words = sub.words;
		sa.stride = sub.stride;
		sa.n_rec = sub.n_rec;
		sa.L.k = k;
		sa.L.ring = ring;
		sa.L.nwarps = nw;
		sa.L.npos_max = npos_max;
		sa.L.start_limit = 0;
		sa.L.mixed_ok = 1; // a candidate on a shorter record's zero padding is dropped by its length (hll_hit_kernel)
		sa.L.prefetch = c->scan_prefetch;
		memcpy(sa.L.F0, c->kinit[0].F0, sizeof sa.L.F0);
		memcpy(sa.L.R0, c->kinit[0].R0, sizeof sa.L.R0);
		sa.masks = c->d_masks;
		sa.tile_info = c->d_tile_info;
		sa.f1_k = c->d_f1;
		sa.cand = reinterpret_cast<unsigned long long*>(c->d_hll_scratch + 2);
		sa.ctl = c->d_hll_scratch + 4;
		sa.hll_min = c->d_hll_scratch;
		sa.grid = std::min<unsigned>((unsigned)c->n_sm, n_tiles);
		sa.smem_bytes = nw * per_warp;
		sa.stream = c->stream;
		if ((rc = stage_begin(c, 0)))
			return rc;
		CK(ntc::pl::launch_hllscan(k, T, sa));
		if ((rc = stage_end(c)) || (rc = stage_begin(c, 1)))
			return rc;
		CK(ntc::launch_hll_hit(sub.words, sub.stride, sub.n_rec, n_tiles, npos_max, c->d_masks, c->d_tile_info, c->d_bs_tab, k, c->hll_bits, c->d_hll,
		    c->n_sm, c->stream));
		if ((rc = stage_end(c)))
			return rc;
		c->n_launches += 2;
		if (T < 13 && (rc = hll_min_request(c))) // the next chunk filters by it
			return rc;
		done += sub.n_rec;
		c->hll_seen += sub.n_rec;
	}
	return NTC_OK;
}

// Long ragged records (long reads, contigs split at N): cut them on the device into equal pieces of Lp bases every D bases
// for the pipeline, plus one ragged tail per record for the general kernel.  Returns the k indices that were handled.
int run_retiled(ntc_ctx* c, const ntc::BatchView& b_in, uint32_t* handled)
{
	*handled = 0;
	if (!c->use_pipeline || c->gap || c->kernel == NTC_KERNEL_ROLL64 || b_in.n_rec == 0 || c->no_retile)
		return NTC_OK;
	if (b_in.n_words / b_in.n_rec < 40) // average record below ~600 bases: pieces would mostly be tails
		return NTC_OK;
	ntc::BatchView b = b_in;
	if (!b.off) { // long records at a uniform stride: give them explicit offsets and treat them like a ragged batch
		if ((uint64_t)b.n_rec * b.stride > 0xFFFFFFF0ull)
			return NTC_OK;
		int rc0;
		if ((rc0 = grow(&c->d_rt_off, &c->cap_rt_off, (size_t)b.n_rec + 1, false)))
			return rc0;
		CK(ntc::launch_stride_offsets(b.stride, b.n_rec, c->d_rt_off, c->stream));
		b.off = c->d_rt_off;
	}
	uint32_t kmask = 0, kmax = 0, kmin = 0xFFFFFFFFu;
	for (unsigned ki = 0; ki < c->nK; ki++)
		if (c->k[ki] <= 160 && ntc::pl::have_scan_kernel(c->k[ki], c->sBits)) {
			kmask |= 1u << ki;
			kmax = std::max(kmax, c->k[ki]);
			kmin = std::min(kmin, c->k[ki]);
		}
	if (!kmask)
		return NTC_OK;
	// piece geometry: at least 4 kmax bases per piece (<= 25 % of the scan spent on overlap), 176 bases (the stride the
	// shared-memory hit kernel takes) when that is enough
	const uint32_t stride = std::max(12u, (1u + (4u * kmax + 15u) / 16u + 3u) & ~3u);
	const uint32_t Lp = 16u * (stride - 1u), D = Lp - kmax + 1u;
	int rc;
	const size_t n1 = (size_t)b.n_rec + 1;
	if ((rc = grow(&c->d_rt_counts, &c->cap_rt_counts, 3 * n1, false)))
		return rc;
	{
		size_t tmp = ntc::piece_scan_temp_bytes(b.n_rec + 1);
		char* t = (char*)c->d_scan_tmp;
		if ((rc = grow(&t, &c->cap_scan_tmp, tmp, false)))
			return rc;
		c->d_scan_tmp = t;
	}
	uint32_t *d_full = c->d_rt_counts, *d_tidx = c->d_rt_counts + n1, *d_tw = c->d_rt_counts + 2 * n1;
	CK(ntc::launch_retile_count(b, Lp, D, kmin, d_full, d_tidx, d_tw, c->d_scan_tmp, c->cap_scan_tmp, c->stream));
	uint32_t tot[3];
	CK(cudaMemcpyAsync(&tot[0], d_full + b.n_rec, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaMemcpyAsync(&tot[1], d_tidx + b.n_rec, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaMemcpyAsync(&tot[2], d_tw + b.n_rec, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	const uint32_t n_full = tot[0], n_tail = tot[1], tail_words = tot[2];
	if ((uint64_t)n_full * stride > 0xFFFFFFF0ull)
		return set_err(NTC_EINVAL, "batch too large to re-tile: %u pieces", n_full);
	if ((rc = grow(&c->d_rt_uniform, &c->cap_rt_uniform, (size_t)n_full * stride + 4, false)) ||
	    (rc = grow(&c->d_rt_tail_words, &c->cap_rt_tail_words, (size_t)tail_words + 4, false)) ||
	    (rc = grow(&c->d_rt_tail_off, &c->cap_rt_tail_off, (size_t)n_tail + 2, false)))
		return rc;
	CK(ntc::launch_retile_fill(b, Lp, D, stride, d_full, d_tidx, d_tw, c->d_rt_uniform, c->d_rt_tail_words, c->d_rt_tail_off, c->n_sm, c->stream));
	c->n_launches += 5;
	c->n_retiled++;
	if (n_full) {
		ntc::BatchView u{ c->d_rt_uniform, nullptr, stride, n_full, (uint64_t)n_full * stride, Lp, 0 };
		PipeShape shape[NTC_MAX_K];
		const uint32_t pm = pipeline_config(c, u, true, shape, /*min_rec=*/1);
		for (unsigned ki = 0; ki < c->nK; ki++)
			if ((kmask >> ki) & 1u) {
				if (!((pm >> ki) & 1u))
					return set_err(NTC_ECUDA, "internal: re-tiled batch not accepted by the pipeline (k index %u)", ki);
				shape[ki].start_limit = D;
			}
		if ((rc = run_pipeline_all(c, u, kmask, shape)))
			return rc;
	}
	if (n_tail) {
		if ((rc = flush(c))) // the general kernel increments the counters in HBM directly
			return rc;
		ntc::BatchView t{ c->d_rt_tail_words, c->d_rt_tail_off, 0, n_tail, tail_words, 0, 0 };
		if ((rc = run_roll64(c, t, false, kmask)))
			return rc;
	}
	*handled = kmask;
	return NTC_OK;
}

// Run the sketch kernels over one device-resident batch on the compute stream.
int run_batch(ntc_ctx* c, const ntc::BatchView& b_in, bool record_is_piece)
{
	if (b_in.n_rec == 0)
		return NTC_OK;
	cudaEvent_t e0, e1;
	int rc;
	if ((rc = get_event(c, &e0)) || (rc = get_event(c, &e1)))
		return rc;
	CK(cudaEventRecord(e0, c->stream));
	if (c->hll_bits) {
		ntc::BatchView hb = b_in;
		if (hb.headerless_len) { // as below: records without length words get one
			const uint32_t s4 = (hb.stride + 1u + 3u) & ~3u;
			if ((uint64_t)hb.n_rec * s4 > 0xFFFFFFF0ull)
				return set_err(NTC_EINVAL, "batch too large: %u records of %u words", hb.n_rec, s4);
			if ((rc = grow(&c->d_rt_uniform, &c->cap_rt_uniform, (size_t)hb.n_rec * s4 + 4, false)))
				return rc;
			CK(ntc::launch_add_headers(hb.words, hb.stride, hb.headerless_len, s4, hb.n_rec, c->d_rt_uniform, c->n_sm, c->stream));
			c->n_launches++;
			hb.words = c->d_rt_uniform;
			hb.stride = s4;
			hb.n_words = (uint64_t)hb.n_rec * s4;
			hb.headerless_len = 0;
		}
		if ((rc = run_hll(c, hb, record_is_piece)))
			return rc;
		CK(cudaEventRecord(e1, c->stream));
		c->timing.emplace_back(e0, e1);
		c->n_batches++;
		return NTC_OK;
	}
	PipeShape shape[NTC_MAX_K];
	ntc::BatchView b = b_in;
	if (b.headerless_len) {
		// records without length words (ntc_submit_bases): give them their length word while padding them to a multiple of 4 words
		const uint32_t s4 = (b.stride + 1u + 3u) & ~3u;
		if ((uint64_t)b.n_rec * s4 > 0xFFFFFFF0ull)
			return set_err(NTC_EINVAL, "batch too large: %u records of %u words", b.n_rec, s4);
		if ((rc = grow(&c->d_rt_uniform, &c->cap_rt_uniform, (size_t)b.n_rec * s4 + 4, false)))
			return rc;
		CK(ntc::launch_add_headers(b.words, b.stride, b.headerless_len, s4, b.n_rec, c->d_rt_uniform, c->n_sm, c->stream));
		c->n_launches++;
		b.words = c->d_rt_uniform;
		b.stride = s4;
		b.n_words = (uint64_t)b.n_rec * s4;
		b.uniform_len = b.headerless_len;
		b.headerless_len = 0;
	}
	if (!b.off && (b.stride & 3u) && b.stride >= 2 && b.n_rec >= 1024 && c->use_pipeline && !c->gap && c->kernel != NTC_KERNEL_ROLL64 &&
	    record_is_piece) {
		// tightly packed uniform batch: pad every record to a multiple of 4 words on the device, then it can take the pipeline
		const uint32_t s4 = (b.stride + 3u) & ~3u;
		bool any = false;
		for (unsigned ki = 0; ki < c->nK; ki++)
			any = any || (c->k[ki] < 288 && ntc::pl::have_scan_kernel(c->k[ki], c->sBits));
		if (any && (uint64_t)b.n_rec * s4 <= 0xFFFFFFF0ull) {
			if ((rc = grow(&c->d_rt_uniform, &c->cap_rt_uniform, (size_t)b.n_rec * s4 + 4, false)))
				return rc;
			CK(ntc::launch_restride(b.words, b.stride, s4, b.n_rec, c->d_rt_uniform, c->n_sm, c->stream));
			c->n_launches++;
			b.words = c->d_rt_uniform;
			b.stride = s4;
			b.n_words = (uint64_t)b.n_rec * s4;
		}
	}
	else if (b.off && b.max_rec_words >= 2 && b.max_rec_words <= 20 && b.n_rec >= 1024 && c->use_pipeline && !c->gap &&
	         (c->kernel == NTC_KERNEL_BITSLICE || (c->kernel == NTC_KERNEL_AUTO && c->pad_ragged)) && record_is_piece) {
		// ragged batch of short records (what ntc_pack_seqs makes of reads: the documented drop-in for ntRead): zero-pad every record
		// to the stride of the longest on the device; the pipeline takes tiles of mixed lengths (scan_kernel.cuh).  Not when the padding
		// would be most of the batch (a few long records among many short fragments).  Halves the device time of such batches, but
		// the host currently spends longer inside ntc_submit on this route than it saves, so NTC_KERNEL_AUTO takes
		// it only with NTC_PAD=1; NTC_KERNEL_BITSLICE (forced pipeline) always does.
		const uint32_t s4 = (b.max_rec_words + 3u) & ~3u;
		const uint64_t padded = (uint64_t)b.n_rec * s4;
		bool any = false;
		for (unsigned ki = 0; ki < c->nK; ki++)
			any = any || (c->k[ki] < 288 && ntc::pl::have_scan_kernel(c->k[ki], c->sBits));
		if (any && padded <= 0xFFFFFFF0ull && padded <= 3 * b.n_words + 4096) {
			if ((rc = grow(&c->d_rt_uniform, &c->cap_rt_uniform, (size_t)padded + 4, false)))
				return rc;
			HT(c, "batch: launch_pad_ragged", CK(ntc::launch_pad_ragged(b.words, b.off, b.n_rec, s4, c->d_rt_uniform, c->n_sm, c->stream)));
			c->n_launches++;
			c->n_padded++;
			b.words = c->d_rt_uniform;
			b.off = nullptr;
			b.stride = s4;
			b.n_words = padded;
		}
	}
	uint32_t pmask = pipeline_config(c, b, record_is_piece, shape);
	uint32_t retiled = 0;
	if (!pmask && (rc = run_retiled(c, b, &retiled)))
		return rc;
	const uint32_t all = c->nK >= 32 ? 0xFFFFFFFFu : ((1u << c->nK) - 1);
	const uint32_t roll_mask = all & ~(pmask | retiled);
	if (c->kernel == NTC_KERNEL_BITSLICE && roll_mask)
		return set_err(NTC_EINVAL, "NTC_KERNEL_BITSLICE forced, but this batch / k / sBits has no bit-sliced variant (kmask %x)", pmask | retiled);
	if (pmask && (rc = run_pipeline_all(c, b, pmask, shape)))
		return rc;
	if (roll_mask && (rc = flush(c))) // the general kernel increments the counters in HBM directly
		return rc;
	if (roll_mask && (rc = run_roll64(c, b, record_is_piece, roll_mask)))
		return rc;
	CK(cudaEventRecord(e1, c->stream));
	c->timing.emplace_back(e0, e1);
	c->n_batches++;
	return NTC_OK;
}

int drain_timing(ntc_ctx* c)
{
	for (auto& pr : c->timing) {
		float ms = 0;
		CK(cudaEventSynchronize(pr.second));
		CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
		c->kernel_ms += ms;
		c->n_timed++;
		c->event_pool.push_back(pr.first);
		c->event_pool.push_back(pr.second);
	}
	c->timing.clear();
	for (auto& sp : c->stage_timing) {
		float ms = 0;
		CK(cudaEventSynchronize(sp.b));
		CK(cudaEventElapsedTime(&ms, sp.a, sp.b));
		c->stage_ms[sp.stage] += ms;
		c->event_pool.push_back(sp.a);
		c->event_pool.push_back(sp.b);
	}
	c->stage_timing.clear();
	return NTC_OK;
}

bool single_piece_records(const ntc_ctx* c, uint32_t max_rec_words)
{
	if (max_rec_words < 1)
		return true;
	uint64_t max_len = (uint64_t)(max_rec_words - 1) * 16;
	return max_len < c->kmin || max_len - c->kmin + 1 <= ntc::PIECE_STARTS;
}

} // namespace

// entry points of the ntCard sketch are not valid on a context made by ntc_hll_create, and vice versa
#define SKETCH_ONLY(c, name)                                                                               \
	do {                                                                                                   \
		if ((c)->hll_bits)                                                                                 \
			return set_err(NTC_ESTATE, name ": not valid on an nthll context (ntc_hll_create)");           \
	} while (0)
#define HLL_ONLY(c, name)                                                                                  \
	do {                                                                                                   \
		if (!(c)->hll_bits)                                                                                \
			return set_err(NTC_ESTATE, name ": needs a context made by ntc_hll_create");                   \
	} while (0)

extern "C" {

const char* ntc_last_error(void) { return ntc::last_err(); }
const char* ntc_version(void) { return "ntcard-b200 0.1 (sm_100a)"; }

int ntc_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

static int create_ctx(ntc_ctx** out, const unsigned* kList, unsigned nK, unsigned rBits, unsigned sBits, int device, void* d_counters,
    void* cuda_stream, unsigned hll_bits, void* d_hll)
{
	if (!out || !kList || nK == 0 || nK > NTC_MAX_K)
		return set_err(NTC_EINVAL, "ntc_create: need 1..%d k values", NTC_MAX_K);
	if (rBits < 1 || rBits > 30 || sBits < 1 || sBits + rBits > 62)
		return set_err(NTC_EINVAL, "ntc_create: rBits=%u sBits=%u out of range", rBits, sBits);
	for (unsigned i = 0; i < nK; i++)
		if (kList[i] == 0)
			return set_err(NTC_EINVAL, "ntc_create: k must be >= 1");
	int ndev = ntc_device_count();
	if (ndev <= 0)
		return set_err(NTC_ENODEVICE, "no CUDA device available (this library has no CPU fallback)");
	if (device < 0 || device >= ndev)
		return set_err(NTC_EINVAL, "ntc_create: device %d of %d", device, ndev);
	ntc_ctx* c = new ntc_ctx();
	c->device = device;
	c->nK = nK;
	c->rBits = rBits;
	c->sBits = sBits;
	c->kmin = c->kmax = kList[0];
	for (unsigned i = 0; i < nK; i++) {
		c->k[i] = kList[i];
		c->kmin = std::min(c->kmin, kList[i]);
		c->kmax = std::max(c->kmax, kList[i]);
	}
	c->n_counters = hll_bits ? 0 : (size_t)nK * NTC_NSAMP << rBits;
	c->hll_bits = hll_bits;
	c->hll_bytes = hll_bits ? std::max((size_t)4, (size_t)1 << hll_bits) : 0;
	int rc = NTC_OK;
	auto fail = [&](int code) {
		ntc_destroy(c);
		return code;
	};
	if ((rc = use_device(c)))
		return fail(rc);
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
		return fail(set_err(NTC_ECUDA, "cudaGetDeviceProperties failed"));
	if (prop.major < 10)
		return fail(set_err(NTC_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor));
	c->n_sm = prop.multiProcessorCount;
#define CKF(call)                                                                                      \
	do {                                                                                               \
		cudaError_t e_ = (call);                                                                       \
		if (e_ != cudaSuccess)                                                                         \
			return fail(set_err(NTC_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)));           \
	} while (0)
	if (cuda_stream) {
		c->stream = (cudaStream_t)cuda_stream;
	} else {
		CKF(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
		c->own_stream = true;
	}
	CKF(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	if (hll_bits) {
		if (d_hll) {
			c->d_hll = (uint8_t*)d_hll;
		} else {
			CKF(cudaMalloc((void**)&c->d_hll, c->hll_bytes));
			c->own_hll = true;
		}
	} else if (d_counters) {
		c->d_counters = (uint32_t*)d_counters;
	} else {
		CKF(cudaMalloc((void**)&c->d_counters, c->n_counters * sizeof(uint32_t)));
		c->own_counters = true;
	}
	CKF(cudaMalloc((void**)&c->d_f1, NTC_MAX_K * sizeof(unsigned long long)));
	CKF(cudaMalloc((void**)&c->d_params, sizeof(ntc::DevParams)));
	ntc::DevParams hp;
	build_params(c, &hp);
	CKF(cudaMemcpy(c->d_params, &hp, sizeof hp, cudaMemcpyHostToDevice));
	{
		std::vector<uint32_t> tab(8 * 256 * 4);
		ntc::pl::build_tables(tab.data());
		CKF(cudaMalloc((void**)&c->d_bs_tab, tab.size() * sizeof(uint32_t)));
		CKF(cudaMemcpy(c->d_bs_tab, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
		for (unsigned ki = 0; ki < nK; ki++) {
			ntc_ctx::KInit& L = c->kinit[ki];
			ntc::bs::init_state(c->k[ki], L.F0, L.R0);
			{
				uint64_t fh = 0, rh = 0; // A^k: NTF64 / NTR64 base forms, nthash.hpp:220-239
				for (unsigned i = 0; i < c->k[ki]; i++) {
					fh = ntc::srol(fh) ^ ntc::seed_of(0);
					rh = ntc::srol(rh) ^ ntc::seed_of(3);
				}
				L.polyA_sampled = !hll_bits && ntc::sample_table(rh < fh ? rh : fh, sBits) < 2;
			}
		}
	}
	c->use_pipeline = !(getenv("NTC_PIPELINE") && atoi(getenv("NTC_PIPELINE")) == 0);
	c->use_fused = getenv("NTC_FUSED") && atoi(getenv("NTC_FUSED")) != 0;
	c->no_stage = getenv("NTC_NO_STAGE") != nullptr;
	if (getenv("NTC_FUSED_DBG"))
		c->fused_dbg = (unsigned)atoi(getenv("NTC_FUSED_DBG"));
	c->no_retile = getenv("NTC_NO_RETILE") != nullptr;
	c->mask_prefetch = getenv("NTC_MASK_PREFETCH") ? (uint32_t)atoi(getenv("NTC_MASK_PREFETCH")) : 1u;
	c->hll_fast = !(getenv("NTC_HLL_FAST") && atoi(getenv("NTC_HLL_FAST")) == 0);
	if (getenv("NTC_HLL_GROW") && atoi(getenv("NTC_HLL_GROW")) >= 10)
		c->hll_grow = (uint32_t)std::min(atoi(getenv("NTC_HLL_GROW")), 10000);
	if (getenv("NTC_HLL_FIRST_K") && atoi(getenv("NTC_HLL_FIRST_K")) >= 4)
		c->hll_first = (uint32_t)std::min(atoi(getenv("NTC_HLL_FIRST_K")), 1 << 20) << 10;
	// two 8-byte cudaMemsetAsync per batch (default) or one 1-thread kernel (NTC_CLEAR_MEMSET=0): measured, the kernel variant makes the host
	// spend ~6 ms per ntc_submit on the ragged host path (tools/bench_ragged.py: 53.7 vs 7.9 ms per pass) -- unexplained, see DESIGN section 8
	c->clear_by_memset = !(getenv("NTC_CLEAR_MEMSET") && atoi(getenv("NTC_CLEAR_MEMSET")) == 0);
	c->host_timing = getenv("NTC_HOST_TIMING") && atoi(getenv("NTC_HOST_TIMING")) != 0;
	c->pad_ragged = !(getenv("NTC_PAD") && atoi(getenv("NTC_PAD")) == 0);
	if (getenv("NTC_SCAN_PREFETCH"))
		c->scan_prefetch = (unsigned)atoi(getenv("NTC_SCAN_PREFETCH"));
	if (getenv("NTC_CHUNK_WAVES"))
		c->chunk_waves = (unsigned)atoi(getenv("NTC_CHUNK_WAVES"));
	if (getenv("NTC_MULTIK_CHUNK_MB"))
		c->multik_chunk_mb = (unsigned)atoi(getenv("NTC_MULTIK_CHUNK_MB"));
	if (!hll_bits && (rc = pool_create(c)))
		return fail(rc);
	for (int i = 0; i < NBUF; i++) {
		CKF(cudaEventCreateWithFlags(&c->stage[i].copied, cudaEventDisableTiming));
		CKF(cudaEventCreateWithFlags(&c->stage[i].consumed, cudaEventDisableTiming));
	}
#undef CKF
	if ((rc = ntc_reset(c)))
		return fail(rc);
	*out = c;
	return NTC_OK;
}

int ntc_create(ntc_ctx** out, const unsigned* kList, unsigned nK, unsigned rBits, unsigned sBits, int device, void* d_counters,
    void* cuda_stream)
{
	return create_ctx(out, kList, nK, rBits, sBits, device, d_counters, cuda_stream, 0, nullptr);
}

int ntc_hll_create(ntc_ctx** out, unsigned k, unsigned nBits, int device, void* d_regs, void* cuda_stream)
{
	if (nBits < 1 || nBits > 30)
		return set_err(NTC_EINVAL, "ntc_hll_create: nBits=%u out of range (1..30)", nBits);
	return create_ctx(out, &k, 1, 1, 1, device, nullptr, cuda_stream, nBits, d_regs);
}

void ntc_destroy(ntc_ctx* c)
{
	if (!c)
		return;
	if (c->host_timing && !c->host_spans.empty()) {
		fprintf(stderr, "[ntc host timing] %-34s %10s %10s %10s\n", "call site", "calls", "total ms", "max ms");
		for (const auto& h : c->host_spans)
			fprintf(stderr, "[ntc host timing] %-34s %10llu %10.3f %10.3f\n", h.what, (unsigned long long)h.n, h.total_ms, h.max_ms);
	}
	cudaSetDevice(c->device);
	if (c->stream)
		cudaStreamSynchronize(c->stream);
	if (c->copy_stream)
		cudaStreamSynchronize(c->copy_stream);
	for (auto& pr : c->timing) {
		cudaEventDestroy(pr.first);
		cudaEventDestroy(pr.second);
	}
	for (auto& sp : c->stage_timing) {
		cudaEventDestroy(sp.a);
		cudaEventDestroy(sp.b);
	}
	for (auto e : c->event_pool)
		cudaEventDestroy(e);
	for (int i = 0; i < NBUF; i++) {
		Stage& s = c->stage[i];
		if (s.d_words) cudaFree(s.d_words);
		if (s.d_off) cudaFree(s.d_off);
		if (s.h_words) cudaFreeHost(s.h_words);
		if (s.h_off) cudaFreeHost(s.h_off);
		if (s.copied) cudaEventDestroy(s.copied);
		if (s.consumed) cudaEventDestroy(s.consumed);
	}
	if (c->d_piece_first) cudaFree(c->d_piece_first);
	if (c->d_piece_rec) cudaFree(c->d_piece_rec);
	if (c->d_scan_tmp) cudaFree(c->d_scan_tmp);
	if (c->d_rt_counts) cudaFree(c->d_rt_counts);
	if (c->d_rt_uniform) cudaFree(c->d_rt_uniform);
	if (c->d_rt_tail_words) cudaFree(c->d_rt_tail_words);
	if (c->d_rt_tail_off) cudaFree(c->d_rt_tail_off);
	if (c->d_rt_off) cudaFree(c->d_rt_off);
	if (c->d_narrow) cudaFree(c->d_narrow);
	if (c->d_phist) cudaFree(c->d_phist);
	if (c->d_f1) cudaFree(c->d_f1);
	if (c->d_params) cudaFree(c->d_params);
	if (c->d_bs_tab) cudaFree(c->d_bs_tab);
	if (c->pool.entries) cudaFree(c->pool.entries);
	for (void* m : c->peer_mapped)
		if (m) cudaIpcCloseMemHandle(m);
	if (c->d_log_slab) cudaFree(c->d_log_slab);
	if (c->pool.gstate) cudaFree(c->pool.gstate);
	if (c->d_masks) cudaFree(c->d_masks);
	for (auto& rs : c->run_slot) {
		if (rs.h) cudaFreeHost(rs.h);
		if (rs.d) cudaFree(rs.d);
		if (rs.done) cudaEventDestroy(rs.done);
	}
	if (c->d_tile_info) cudaFree(c->d_tile_info);
	if (c->d_offchk) cudaFree(c->d_offchk);
	if (c->d_hll_scratch) cudaFree(c->d_hll_scratch);
	if (c->h_hll_min) cudaFreeHost(c->h_hll_min);
	if (c->hll_min_ev) cudaEventDestroy(c->hll_min_ev);
	if (c->own_counters && c->d_counters) cudaFree(c->d_counters);
	if (c->own_hll && c->d_hll) cudaFree(c->d_hll);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

int ntc_reset(ntc_ctx* c)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	if (c->hll_bits) { // nthll mode: fresh registers (nthll.cpp:200-201)
		CK(cudaMemsetAsync(c->d_hll, 0, c->hll_bytes, c->stream));
		CK(cudaMemsetAsync(c->d_f1, 0, NTC_MAX_K * sizeof(unsigned long long), c->stream));
		if (c->hll_min_pending) // a minimum still in flight belongs to the old registers
			CK(cudaEventSynchronize(c->hll_min_ev));
		c->hll_min_pending = false;
		c->hll_min = 0;
		c->hll_seen = 0;
		c->totals_overridden = false;
		c->pending = false;
		return NTC_OK;
	}
	// The counters are NOT cleared here: the first flush after a reset writes zeros slice by slice right before it
	// applies the slice's increments (apply_kernel, state 0), which saves one full pass over the 1 GiB/k sketch.
	CK(cudaMemsetAsync(c->d_pool_ctl_region, 0, c->pool_ctl_bytes, c->stream)); // empty log, state = not materialised
	CK(cudaMemsetAsync(c->d_f1, 0, NTC_MAX_K * sizeof(unsigned long long), c->stream));
	{ // the hit groups' saved open / spare blocks belong to the old log: invalidate their generation words
		const ntc::pl::Pool& P = c->pool;
		const size_t pitch = ntc::pl::gstate_row(P.nbins) * sizeof(uint32_t);
		CK(cudaMemset2DAsync(P.gstate, pitch, 0, 2 * sizeof(uint32_t), (size_t)c->nK * P.max_groups, c->stream));
	}
	c->pool.epoch++; // (and date the new ones differently)
	if (c->pool.epoch == 0)
		c->pool.epoch = 1;
	c->totals_overridden = false;
	c->pending = true;
	c->partial = false;
	c->h_nblk.clear();
	return NTC_OK;
}

int ntc_set_gap(ntc_ctx* c, unsigned gap)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	SKETCH_ONLY(c, "ntc_set_gap");
	if (gap != 0) {
		if (c->nK != 1)
			return set_err(NTC_EINVAL, "ntc_set_gap: gap seeds support one k only (ntcard.cpp:397)");
		if (gap % 2 != c->k[0] % 2)
			return set_err(NTC_EINVAL, "ntc_set_gap: gap size and k must have the same modulus (ntcard.cpp:382)");
		if (gap + 2 > c->k[0])
			return set_err(NTC_EINVAL, "ntc_set_gap: gap %u does not fit k = %u", gap, c->k[0]);
	}
	int rc;
	if ((rc = ntc_sync(c))) // batches already submitted keep the old seed
		return rc;
	c->gap = gap;
	ntc::DevParams hp;
	build_params(c, &hp);
	CK(cudaMemcpy(c->d_params, &hp, sizeof hp, cudaMemcpyHostToDevice));
	return NTC_OK;
}

int ntc_set_kernel(ntc_ctx* c, int kernel)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || kernel < NTC_KERNEL_AUTO || kernel > NTC_KERNEL_BITSLICE)
		return set_err(NTC_EINVAL, "ntc_set_kernel: bad argument");
	c->kernel = kernel;
	return NTC_OK;
}

static int submit_host(ntc_ctx* c, const uint32_t* words, size_t n_words, const uint32_t* off, size_t n_rec, uint32_t stride_words,
    uint32_t headerless_len, uint64_t* ticket);

int ntc_submit(ntc_ctx* c, const uint32_t* words, size_t n_words, const uint32_t* off, size_t n_rec, uint32_t stride_words,
    uint64_t* ticket)
{
	return submit_host(c, words, n_words, off, n_rec, stride_words, 0, ticket);
}

int ntc_submit_bases(ntc_ctx* c, const uint32_t* bases, size_t n_rec, uint32_t len_bases, uint64_t* ticket)
{
	if (len_bases == 0)
		return set_err(NTC_EINVAL, "ntc_submit_bases: len_bases must be >= 1");
	const uint32_t wpr = (len_bases + 15u) / 16u;
	return submit_host(c, bases, n_rec * (size_t)wpr, nullptr, n_rec, wpr, len_bases, ticket);
}

static int submit_host(ntc_ctx* c, const uint32_t* words, size_t n_words, const uint32_t* off, size_t n_rec, uint32_t stride_words,
    uint32_t headerless_len, uint64_t* ticket)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || (!words && n_words) || n_rec > 0xFFFFFFF0ull || n_words > 0xFFFFFFF0ull)
		return set_err(NTC_EINVAL, "ntc_submit: bad argument");
	if (!off && n_rec && (stride_words == 0 || (uint64_t)stride_words * n_rec > n_words))
		return set_err(NTC_EINVAL, "ntc_submit: uniform batch needs stride_words*n_rec <= n_words");
	if (ticket)
		*ticket = 0;
	if (n_rec == 0)
		return NTC_OK;
	if (c->partial)
		return set_err(NTC_ESTATE, "ntc_submit: the sketch was flushed partially (ntc_flush_slices); ntc_reset first");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	uint32_t max_rec_words = stride_words;
	if (off) {
		const auto v0 = std::chrono::steady_clock::now();
		if ((rc = ntc_check_offsets(off, n_rec, n_words, &max_rec_words)))
			return rc;
		if (c->host_timing)
			c->host_span("submit: offset validation loop", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - v0).count());
	}
	c->h_nblk.clear(); // a snapshot taken by ntc_log_counts is stale once anything more is logged
	Stage& s = c->stage[c->next_ticket % NBUF];
	// the slot's previous batch must have been consumed by its kernels before we overwrite it
	HT(c, "submit: wait for the slot (consumed)", CK(cudaEventSynchronize(s.consumed)));
	if ((rc = grow(&s.d_words, &s.cap_words, n_words, false)))
		return rc;
	if (off && (rc = grow(&s.d_off, &s.cap_off, n_rec + 1, false)))
		return rc;
	cudaPointerAttributes attr;
	bool pinned = cudaPointerGetAttributes(&attr, words) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	const uint32_t* src_words = words;
	const uint32_t* src_off = off;
	if (!pinned) {
		CK(cudaEventSynchronize(s.copied)); // staging of this slot is free again
		if ((rc = grow(&s.h_words, &s.cap_h_words, n_words, true)))
			return rc;
		memcpy(s.h_words, words, n_words * sizeof(uint32_t));
		src_words = s.h_words;
	}
	if (off) {
		cudaPointerAttributes a2;
		bool pinned_off = cudaPointerGetAttributes(&a2, off) == cudaSuccess && a2.type == cudaMemoryTypeHost;
		cudaGetLastError();
		if (!pinned_off) {
			if (pinned)
				CK(cudaEventSynchronize(s.copied));
			if ((rc = grow(&s.h_off, &s.cap_h_off, n_rec + 1, true)))
				return rc;
			memcpy(s.h_off, off, (n_rec + 1) * sizeof(uint32_t));
			src_off = s.h_off;
		}
	}
	HT(c, "submit: cudaMemcpyAsync words", CK(cudaMemcpyAsync(s.d_words, src_words, n_words * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream)));
	if (off)
		HT(c, "submit: cudaMemcpyAsync offsets", CK(cudaMemcpyAsync(s.d_off, src_off, (n_rec + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream)));
	HT(c, "submit: record copied + stream wait", CK(cudaEventRecord(s.copied, c->copy_stream)); CK(cudaStreamWaitEvent(c->stream, s.copied, 0)));
	ntc::BatchView b{ s.d_words, off ? s.d_off : nullptr, stride_words, (uint32_t)n_rec, n_words, 0, off ? max_rec_words : 0u };
	b.headerless_len = headerless_len;
	HT(c, "submit: run_batch (all of it)", rc = run_batch(c, b, single_piece_records(c, headerless_len ? max_rec_words + 1 : max_rec_words)));
	if (rc)
		return rc;
	CK(cudaEventRecord(s.consumed, c->stream));
	s.ticket = c->next_ticket++;
	if (ticket)
		*ticket = s.ticket;
	return NTC_OK;
}

int ntc_submit_device(ntc_ctx* c, const uint32_t* d_words, size_t n_words, const uint32_t* d_off, size_t n_rec,
    uint32_t stride_words)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || (!d_words && n_words) || n_rec > 0xFFFFFFF0ull || n_words > 0xFFFFFFF0ull)
		return set_err(NTC_EINVAL, "ntc_submit_device: bad argument");
	if (!d_off && n_rec && (stride_words == 0 || (uint64_t)stride_words * n_rec > n_words))
		return set_err(NTC_EINVAL, "ntc_submit_device: uniform batch needs stride_words*n_rec <= n_words");
	if (n_rec == 0)
		return NTC_OK;
	if (c->partial)
		return set_err(NTC_ESTATE, "ntc_submit_device: the sketch was flushed partially (ntc_flush_slices); ntc_reset first");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	c->h_nblk.clear(); // a snapshot taken by ntc_log_counts is stale once anything more is logged
	uint32_t max_rec_words = 0;
	if (d_off) {
		// what ntc_check_offsets does for host batches, on the device: validity of the offsets and the longest record (the
		// host needs it to choose the kernels: short records are padded for the pipeline, long ones re-tiled)
		if (!c->d_offchk)
			CK(cudaMalloc((void**)&c->d_offchk, 2 * sizeof(uint32_t)));
		uint32_t h[2] = { 0, 0 };
		CK(cudaMemsetAsync(c->d_offchk, 0, sizeof h, c->stream));
		CK(ntc::launch_check_offsets(d_off, (uint32_t)n_rec, n_words, c->d_offchk, c->n_sm, c->stream));
		CK(cudaMemcpyAsync(h, c->d_offchk, sizeof h, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		c->n_launches++;
		if (h[1])
			return set_err(NTC_EINVAL, "ntc_submit_device: offsets must start at 0, be non-decreasing and end within n_words");
		max_rec_words = h[0];
	}
	ntc::BatchView b{ d_words, d_off, stride_words, (uint32_t)n_rec, n_words, 0, max_rec_words };
	return run_batch(c, b, d_off ? single_piece_records(c, max_rec_words) : single_piece_records(c, stride_words));
}

int ntc_wait(ntc_ctx* c, uint64_t ticket)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	if (ticket == 0)
		return NTC_OK;
	for (int i = 0; i < NBUF; i++)
		if (c->stage[i].ticket == ticket) {
			CK(cudaEventSynchronize(c->stage[i].copied));
			return NTC_OK;
		}
	return NTC_OK; // older than the ring: long finished
}

int ntc_sync(ntc_ctx* c)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	if ((rc = flush(c)))
		return rc;
	CK(cudaStreamSynchronize(c->copy_stream));
	CK(cudaStreamSynchronize(c->stream));
	return drain_timing(c);
}

int ntc_flush(ntc_ctx* c)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	return flush(c);
}


/* ---- hit-log exchange (multi-GPU sparse reduction) ------------------------------------------------------ */
int ntc_log_info(ntc_ctx* c, uint32_t* n_slices, uint64_t* counters_per_slice, uint32_t* entries_per_block)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	SKETCH_ONLY(c, "ntc_log_info");
	if (n_slices) *n_slices = c->pool.n_slices;
	if (counters_per_slice) *counters_per_slice = (uint64_t)1 << c->pool.bin_shift;
	if (entries_per_block) *entries_per_block = ntc::pl::kBlkEntries;
	return NTC_OK;
}

int ntc_log_counts(ntc_ctx* c, uint32_t* nblk, int* exportable, uint32_t* pool_info)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !nblk || !exportable)
		return set_err(NTC_EINVAL, "ntc_log_counts: bad argument");
	SKETCH_ONLY(c, "ntc_log_counts");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	const ntc::pl::Pool& P = c->pool;
	std::vector<uint32_t> h(ntc::pl::CTL_WORDS + P.n_slices);
	CK(cudaMemcpyAsync(h.data(), P.ctl, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream)); // ctl and slice_nblk are adjacent
	CK(cudaStreamSynchronize(c->stream));
	if ((rc = drain_timing(c)))
		return rc;
	c->h_nblk.assign(P.n_slices, 0);
	for (uint32_t s = 0; s < P.n_slices; s++)
		nblk[s] = c->h_nblk[s] = std::min(h[ntc::pl::CTL_WORDS + s], P.slice_cap);
	c->log_used = std::min(h[ntc::pl::CTL_NEXT], P.n_blocks);
	if (pool_info) {
		pool_info[0] = c->log_used;
		pool_info[1] = P.n_blocks;
		pool_info[2] = P.slice_cap;
	}
	// exportable: everything submitted since the reset is still in the log (nothing flushed, nothing added directly)
	*exportable = (h[ntc::pl::CTL_STATE] == 0 && h[ntc::pl::CTL_DIRECT] == 0 && h[ntc::pl::CTL_NEXT] <= P.n_blocks && c->pending && !c->partial) ? 1 : 0;
	return NTC_OK;
}

// Small host tables (runs of slices, slice orders) for the kernels of the exchange / partial flush: staged in a ring of
// pinned buffers and copied asynchronously, so the calls do not drain the stream.  Each slot has its own device copy.
static int upload_runs(ntc_ctx* c, const std::vector<uint32_t>& runs, const uint32_t** d_out)
{
	ntc_ctx::RunSlot& s = c->run_slot[c->next_run_slot++ % ntc_ctx::kRunSlots];
	int rc;
	if (!s.done)
		CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
	CK(cudaEventSynchronize(s.done)); // the slot's previous copy has executed (long ago)
	if ((rc = grow(&s.h, &s.cap_h, runs.size() + 2, true)) || (rc = grow(&s.d, &s.cap_d, runs.size() + 2, false)))
		return rc;
	memcpy(s.h, runs.data(), runs.size() * sizeof(uint32_t));
	CK(cudaMemcpyAsync(s.d, s.h, runs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
	CK(cudaEventRecord(s.done, c->stream));
	*d_out = s.d;
	return NTC_OK;
}

int ntc_log_export(ntc_ctx* c, const uint32_t* slices, uint32_t n, void* d_blocks)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || (n && !slices) || c->h_nblk.size() != c->pool.n_slices)
		return set_err(c && c->h_nblk.empty() ? NTC_ESTATE : NTC_EINVAL, "ntc_log_export: bad argument, or no current ntc_log_counts snapshot (a batch, flush or reset since invalidates it)");
	SKETCH_ONLY(c, "ntc_log_export");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	std::vector<uint32_t> runs;
	uint32_t total = 0;
	for (uint32_t i = 0; i < n; i++) {
		if (slices[i] >= c->pool.n_slices)
			return set_err(NTC_EINVAL, "ntc_log_export: slice %u of %u", slices[i], c->pool.n_slices);
		if (c->h_nblk[slices[i]] == 0)
			continue;
		runs.push_back(slices[i]);
		runs.push_back(total);
		total += c->h_nblk[slices[i]];
	}
	if (total == 0)
		return NTC_OK;
	if (!d_blocks)
		return set_err(NTC_EINVAL, "ntc_log_export: null output buffer");
	const uint32_t* d_runs = nullptr;
	if ((rc = upload_runs(c, runs, &d_runs)))
		return rc;
	CK(ntc::pl::launch_export(c->pool, d_runs, (uint32_t)runs.size() / 2, total, (uint32_t*)d_blocks, c->stream));
	c->n_launches++;
	return NTC_OK;
}

int ntc_log_import(ntc_ctx* c, const void* d_blocks, uint32_t n_blocks, const uint32_t* runs_in, uint32_t n_runs)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || (n_blocks && (!d_blocks || !runs_in)) || c->h_nblk.size() != c->pool.n_slices)
		return set_err(c && c->h_nblk.empty() ? NTC_ESTATE : NTC_EINVAL, "ntc_log_import: bad argument, or no current ntc_log_counts snapshot (a batch, flush or reset since invalidates it)");
	SKETCH_ONLY(c, "ntc_log_import");
	if (n_blocks == 0)
		return NTC_OK;
	int rc;
	if ((rc = use_device(c)))
		return rc;
	const ntc::pl::Pool& P = c->pool;
	if ((uint64_t)c->log_used + n_blocks > P.n_blocks)
		return set_err(NTC_ENOMEM, "ntc_log_import: %u blocks do not fit the hit log (%u of %u used)", n_blocks, c->log_used, P.n_blocks);
	std::vector<uint32_t> runs;
	uint32_t total = 0;
	for (uint32_t i = 0; i < n_runs; i++) {
		const uint32_t s = runs_in[2 * i], nb = runs_in[2 * i + 1];
		if (s >= P.n_slices)
			return set_err(NTC_EINVAL, "ntc_log_import: slice %u of %u", s, P.n_slices);
		if (nb == 0)
			continue;
		if ((uint64_t)c->h_nblk[s] + nb > P.slice_cap)
			return set_err(NTC_ENOMEM, "ntc_log_import: block list of slice %u is full", s);
		c->h_nblk[s] += nb;
		runs.push_back(s);
		runs.push_back(total);
		total += nb;
	}
	if (total != n_blocks)
		return set_err(NTC_EINVAL, "ntc_log_import: runs cover %u blocks, n_blocks = %u", total, n_blocks);
	const uint32_t* d_runs = nullptr;
	if ((rc = upload_runs(c, runs, &d_runs)))
		return rc;
	// wire blocks (260 words) -> pool blocks (256 words)
	CK(cudaMemcpy2DAsync(P.entries + (size_t)c->log_used * ntc::pl::kBlkEntries, ntc::pl::kBlkEntries * sizeof(uint32_t), d_blocks,
	    ntc::pl::kWireBlkWords * sizeof(uint32_t), ntc::pl::kBlkEntries * sizeof(uint32_t), n_blocks, cudaMemcpyDeviceToDevice, c->stream));
	CK(ntc::pl::launch_import(P, d_runs, (uint32_t)runs.size() / 2, n_blocks, c->log_used, (const uint32_t*)d_blocks, c->stream));
	c->log_used += n_blocks;
	c->n_launches++;
	c->pending = true;
	return NTC_OK;
}

/* ---- multi-GPU reduction over peer memory (one process per GPU, one node) ------------------------------------------ */
int ntc_peer_export(ntc_ctx* c, void* handles)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !handles)
		return set_err(NTC_EINVAL, "ntc_peer_export: bad argument");
	SKETCH_ONLY(c, "ntc_peer_export");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	cudaIpcMemHandle_t h[2];
	CK(cudaIpcGetMemHandle(&h[0], c->pool.entries));
	CK(cudaIpcGetMemHandle(&h[1], c->d_log_slab));
	static_assert(sizeof(cudaIpcMemHandle_t) == NTC_PEER_HANDLE_BYTES / 2, "CUDA IPC handle size");
	memcpy(handles, h, sizeof h);
	return NTC_OK;
}

int ntc_peer_attach(ntc_ctx* c, int world, int rank, const void* all_handles)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !all_handles || world < 1 || world > (int)ntc::pl::kMaxPeers || rank < 0 || rank >= world)
		return set_err(NTC_EINVAL, "ntc_peer_attach: bad argument (world 1..%u)", ntc::pl::kMaxPeers);
	SKETCH_ONLY(c, "ntc_peer_attach");
	if (c->peer_world)
		return set_err(NTC_ESTATE, "ntc_peer_attach: already attached");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	const ntc::pl::Pool& P = c->pool;
	for (int q = 0; q < world; q++) {
		const uint32_t *ent, *slab;
		if (q == rank) {
			ent = P.entries;
			slab = c->d_log_slab;
		} else {
			cudaIpcMemHandle_t h[2];
			memcpy(h, (const char*)all_handles + (size_t)q * NTC_PEER_HANDLE_BYTES, sizeof h);
			void *m0 = nullptr, *m1 = nullptr;
			CK(cudaIpcOpenMemHandle(&m0, h[0], cudaIpcMemLazyEnablePeerAccess));
			c->peer_mapped[2 * q] = m0;
			CK(cudaIpcOpenMemHandle(&m1, h[1], cudaIpcMemLazyEnablePeerAccess));
			c->peer_mapped[2 * q + 1] = m1;
			ent = (const uint32_t*)m0;
			slab = (const uint32_t*)m1;
		}
		c->peers.entries[q] = ent;
		c->peers.slice_blocks[q] = slab + c->log_slab_ctl_words;
		c->peers.slice_nblk[q] = slab + ntc::pl::CTL_WORDS;
	}
	c->peers.n = (uint32_t)world;
	c->peers.self = (uint32_t)rank;
	c->peer_world = world;
	c->peer_rank = rank;
	return NTC_OK;
}

int ntc_peer_attach_contexts(ntc_ctx* c, int world, int rank, ntc_ctx* const* all)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !all || world < 1 || world > (int)ntc::pl::kMaxPeers || rank < 0 || rank >= world || all[rank] != c)
		return set_err(NTC_EINVAL, "ntc_peer_attach_contexts: bad argument (world 1..%u, all[rank] must be the context itself)", ntc::pl::kMaxPeers);
	SKETCH_ONLY(c, "ntc_peer_attach_contexts");
	if (c->peer_world)
		return set_err(NTC_ESTATE, "ntc_peer_attach_contexts: already attached");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	for (int q = 0; q < world; q++) {
		const ntc_ctx* o = all[q];
		if (!o || o->hll_bits || o->nK != c->nK || o->rBits != c->rBits || o->pool.n_blocks != c->pool.n_blocks || o->pool.nbins != c->pool.nbins)
			return set_err(NTC_EINVAL, "ntc_peer_attach_contexts: context %d has a different geometry", q);
		if (o->device != c->device) {
			int can = 0;
			CK(cudaDeviceCanAccessPeer(&can, c->device, o->device));
			if (!can)
				return set_err(NTC_ECUDA, "ntc_peer_attach_contexts: device %d cannot access device %d", c->device, o->device);
			cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
				CK(e);
			cudaGetLastError();
		}
		c->peers.entries[q] = o->pool.entries;
		c->peers.slice_blocks[q] = o->pool.slice_blocks;
		c->peers.slice_nblk[q] = o->pool.slice_nblk;
	}
	c->peers.n = (uint32_t)world;
	c->peers.self = (uint32_t)rank;
	c->peer_world = world;
	c->peer_rank = rank;
	return NTC_OK;
}

int ntc_log_status_device(ntc_ctx* c, void* d_status)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !d_status)
		return set_err(NTC_EINVAL, "ntc_log_status_device: bad argument");
	SKETCH_ONLY(c, "ntc_log_status_device");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	CK(ntc::pl::launch_log_status(c->pool, c->d_f1, c->nK, (c->pending && !c->partial) ? 1 : 0, (long long*)d_status, c->stream));
	c->n_launches++;
	return NTC_OK;
}

int ntc_reduce_owned(ntc_ctx* c, const void* d_status, void* d_p_hist)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !d_p_hist)
		return set_err(NTC_EINVAL, "ntc_reduce_owned: bad argument");
	SKETCH_ONLY(c, "ntc_reduce_owned");
	if (!c->peer_world)
		return set_err(NTC_ESTATE, "ntc_reduce_owned: ntc_peer_attach first");
	if (c->partial)
		return set_err(NTC_ESTATE, "ntc_reduce_owned: the sketch was already flushed partially; ntc_reset first");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	const ntc::pl::Pool& P = c->pool;
	std::vector<uint32_t> order;
	for (uint32_t s = 0; s < P.n_slices; s++)
		if ((int)(s % (uint32_t)c->peer_world) == c->peer_rank)
			order.push_back(s);
	const uint32_t n = (uint32_t)order.size();
	order.push_back(0);
	const uint32_t* d_order = nullptr;
	if ((rc = upload_runs(c, order, &d_order)))
		return rc;
	const size_t hist_bytes = (size_t)c->nK * NTC_NSAMP * 65536 * sizeof(uint32_t);
	cudaEvent_t e0, e1;
	if ((rc = get_event(c, &e0)) || (rc = get_event(c, &e1)))
		return rc;
	CK(cudaEventRecord(e0, c->stream));
	if ((rc = stage_begin(c, 2)))
		return rc;
	CK(cudaMemsetAsync(d_p_hist, 0, hist_bytes, c->stream));
	CK(ntc::pl::launch_apply_owned(P, c->peers, c->d_counters, d_order, n, (const long long*)d_status, c->nK, c->apply_grid, c->stream));
	cudaError_t e = ntc::pl::launch_hist_slices(P, c->d_counters, d_order, n, (uint32_t*)d_p_hist, c->stream);
	if (e == cudaErrorInvalidValue)
		return set_err(NTC_EINVAL, "ntc_reduce_owned: rBits too small for the histogram chunk");
	CK(e);
	if ((rc = stage_end(c)))
		return rc;
	CK(cudaEventRecord(e1, c->stream));
	c->timing.emplace_back(e0, e1);
	c->n_launches += 2;
	c->pending = false;
	c->partial = true;
	return NTC_OK;
}

int ntc_stream_sync(ntc_ctx* c)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	CK(cudaStreamSynchronize(c->stream));
	return NTC_OK;
}

int ntc_flush_slices(ntc_ctx* c, const uint8_t* owned)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !owned)
		return set_err(NTC_EINVAL, "ntc_flush_slices: bad argument");
	SKETCH_ONLY(c, "ntc_flush_slices");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	std::vector<uint32_t> order;
	for (uint32_t s = 0; s < c->pool.n_slices; s++)
		if (owned[s])
			order.push_back(s);
	const uint32_t n = (uint32_t)order.size();
	order.push_back(0); // never an empty upload
	const uint32_t* d_order = nullptr;
	if ((rc = upload_runs(c, order, &d_order)))
		return rc;
	cudaEvent_t e0, e1;
	if ((rc = get_event(c, &e0)) || (rc = get_event(c, &e1)))
		return rc;
	CK(cudaEventRecord(e0, c->stream));
	if ((rc = stage_begin(c, 2)))
		return rc;
	CK(ntc::pl::launch_apply(c->pool, c->d_counters, 1, 0, c->apply_grid, c->stream, d_order, n));
	if ((rc = stage_end(c)))
		return rc;
	CK(cudaEventRecord(e1, c->stream));
	c->timing.emplace_back(e0, e1);
	c->n_launches++;
	c->pending = false;
	c->partial = true;
	return NTC_OK;
}

int ntc_hist_slices(ntc_ctx* c, const uint8_t* owned, uint32_t* p_hist, void* d_p_hist)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !owned || (!p_hist && !d_p_hist))
		return set_err(NTC_EINVAL, "ntc_hist_slices: bad argument");
	SKETCH_ONLY(c, "ntc_hist_slices");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	if (c->pending)
		return set_err(NTC_ESTATE, "ntc_hist_slices: flush first (ntc_flush_slices / ntc_flush)");
	const ntc::pl::Pool& P = c->pool;
	const uint32_t n_tables = c->nK * NTC_NSAMP;
	const size_t hist_bytes = (size_t)n_tables * 65536 * sizeof(uint32_t);
	uint32_t* dst = (uint32_t*)d_p_hist;
	if (!dst) {
		if (!c->d_phist)
			CK(cudaMalloc((void**)&c->d_phist, hist_bytes));
		dst = c->d_phist;
	}
	CK(cudaMemsetAsync(dst, 0, hist_bytes, c->stream));
	std::vector<uint32_t> order;
	for (uint32_t s = 0; s < P.n_slices; s++)
		if (owned[s])
			order.push_back(s);
	const uint32_t n = (uint32_t)order.size();
	order.push_back(0);
	const uint32_t* d_order = nullptr;
	if ((rc = upload_runs(c, order, &d_order)))
		return rc;
	cudaError_t e = ntc::pl::launch_hist_slices(P, c->d_counters, d_order, n, dst, c->stream);
	if (e == cudaErrorInvalidValue)
		return set_err(NTC_EINVAL, "ntc_hist_slices: rBits too small for the histogram chunk");
	CK(e);
	c->n_launches++;
	if (p_hist) {
		CK(cudaMemcpyAsync(p_hist, dst, hist_bytes, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return NTC_OK;
}

int ntc_counters_device(ntc_ctx* c, void** d_counters, size_t* n_counters)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	SKETCH_ONLY(c, "ntc_counters_device");
	int rc;
	if (c->partial)
		return set_err(NTC_ESTATE, "ntc_counters_device: after ntc_flush_slices only ntc_hist_slices / ntc_totals / ntc_reset are valid");
	if ((rc = use_device(c)) || (rc = flush(c))) // what the caller reads (or all-reduces) must be complete
		return rc;
	if (d_counters) *d_counters = c->d_counters;
	if (n_counters) *n_counters = c->n_counters;
	return NTC_OK;
}

int ntc_totals(ntc_ctx* c, uint64_t* totKmer)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !totKmer)
		return set_err(NTC_EINVAL, "ntc_totals: bad argument");
	if (c->totals_overridden) {
		memcpy(totKmer, c->totals, c->nK * sizeof(uint64_t));
		return NTC_OK;
	}
	int rc;
	if ((rc = ntc_sync(c)))
		return rc;
	unsigned long long h[NTC_MAX_K];
	CK(cudaMemcpy(h, c->d_f1, sizeof h, cudaMemcpyDeviceToHost));
	for (unsigned i = 0; i < c->nK; i++)
		totKmer[i] = h[i];
	return NTC_OK;
}

int ntc_totals_nosync(ntc_ctx* c, uint64_t* totKmer)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !totKmer)
		return set_err(NTC_EINVAL, "ntc_totals_nosync: bad argument");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	unsigned long long h[NTC_MAX_K];
	CK(cudaMemcpyAsync(h, c->d_f1, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	for (unsigned i = 0; i < c->nK; i++)
		totKmer[i] = h[i];
	return NTC_OK;
}

int ntc_set_totals(ntc_ctx* c, const uint64_t* totKmer)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !totKmer)
		return set_err(NTC_EINVAL, "ntc_set_totals: bad argument");
	memcpy(c->totals, totKmer, c->nK * sizeof(uint64_t));
	c->totals_overridden = true;
	return NTC_OK;
}

int ntc_finish(ntc_ctx* c, uint16_t* t_Counter, uint64_t* totKmer, uint32_t* p_hist)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	SKETCH_ONLY(c, "ntc_finish");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	if (c->partial)
		return set_err(NTC_ESTATE, "ntc_finish: after ntc_flush_slices only ntc_hist_slices / ntc_totals / ntc_reset are valid");
	if ((rc = flush(c)))
		return rc;
	if (totKmer && (rc = ntc_totals(c, totKmer)))
		return rc;
	const uint32_t n_tables = c->nK * NTC_NSAMP;
	const uint64_t n_per_table = (uint64_t)1 << c->rBits;
	if (t_Counter && !c->d_narrow)
		CK(cudaMalloc((void**)&c->d_narrow, c->n_counters * sizeof(uint16_t)));
	if (p_hist) {
		if (!c->d_phist)
			CK(cudaMalloc((void**)&c->d_phist, (size_t)n_tables * 65536 * sizeof(uint32_t)));
		CK(cudaMemsetAsync(c->d_phist, 0, (size_t)n_tables * 65536 * sizeof(uint32_t), c->stream));
	}
	if (!t_Counter && p_hist && ntc::pl::launch_hist_slices(c->pool, c->d_counters, nullptr, 0, c->d_phist, c->stream) == cudaSuccess) {
		c->n_launches++; // histogram only: the vectorised kernel, slice by slice
	} else if (t_Counter || p_hist) {
		cudaGetLastError();
		CK(ntc::launch_narrow_hist(c->d_counters, n_tables, n_per_table, t_Counter ? c->d_narrow : nullptr,
		    p_hist ? c->d_phist : nullptr, c->stream));
		c->n_launches++;
	}
	if (t_Counter)
		CK(cudaMemcpyAsync(t_Counter, c->d_narrow, c->n_counters * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
	if (p_hist)
		CK(cudaMemcpyAsync(p_hist, c->d_phist, (size_t)n_tables * 65536 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	if ((rc = ntc_sync(c)))
		return rc;
	if (p_hist) {
		// the kernel counts values >= 1 only; p[0] is what is left of the 2^rBits buckets
		for (uint32_t t = 0; t < n_tables; t++) {
			uint64_t nz = 0;
			uint32_t* p = p_hist + (size_t)t * 65536;
			for (uint32_t v = 1; v < 65536; v++)
				nz += p[v];
			p[0] = (uint32_t)(n_per_table - nz);
		}
	}
	return NTC_OK;
}

// ---- nthll mode --------------------------------------------------------------------------------------------
int ntc_hll_registers_device(ntc_ctx* c, void** d_regs, size_t* n_regs)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	HLL_ONLY(c, "ntc_hll_registers_device");
	if (d_regs) *d_regs = c->d_hll;
	if (n_regs) *n_regs = (size_t)1 << c->hll_bits;
	return NTC_OK;
}

int ntc_hll_finish(ntc_ctx* c, uint8_t* regs, uint64_t* totKmer)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	HLL_ONLY(c, "ntc_hll_finish");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	if (regs)
		CK(cudaMemcpyAsync(regs, c->d_hll, (size_t)1 << c->hll_bits, cudaMemcpyDeviceToHost, c->stream));
	if ((rc = ntc_sync(c)))
		return rc;
	if (totKmer && (rc = ntc_totals(c, totKmer)))
		return rc;
	return NTC_OK;
}

int ntc_hist_range(ntc_ctx* c, const void* d_counters, uint64_t first, uint64_t n, uint32_t* p_hist)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !d_counters || !p_hist || first + n > c->n_counters)
		return set_err(NTC_EINVAL, "ntc_hist_range: bad argument");
	SKETCH_ONLY(c, "ntc_hist_range");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	if ((rc = flush(c)))
		return rc;
	const uint32_t n_tables = c->nK * NTC_NSAMP;
	const size_t hist_bytes = (size_t)n_tables * 65536 * sizeof(uint32_t);
	if (!c->d_phist)
		CK(cudaMalloc((void**)&c->d_phist, hist_bytes));
	CK(cudaMemsetAsync(c->d_phist, 0, hist_bytes, c->stream));
	cudaError_t e = ntc::launch_hist_range((const uint32_t*)d_counters, first, n, c->rBits, c->d_phist, c->stream);
	if (e == cudaErrorInvalidValue)
		return set_err(NTC_EINVAL, "ntc_hist_range: first and n must be multiples of min(65536, 2^rBits)");
	CK(e);
	c->n_launches++;
	CK(cudaMemcpyAsync(p_hist, c->d_phist, hist_bytes, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return NTC_OK;
}

void* ntc_host_alloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
		set_err(NTC_ENOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
		return nullptr;
	}
	return p;
}

void ntc_host_free(void* p)
{
	if (p)
		cudaFreeHost(p);
}

int ntc_gen_packed_device(ntc_ctx* c, uint64_t seed, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U,
    uint32_t stride_words, uint32_t* d_words)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !d_words || stride_words < 1 + (L + 15) / 16 || mode < 0 || mode > 1)
		return set_err(NTC_EINVAL, "ntc_gen_packed_device: bad argument");
	int rc;
	if ((rc = use_device(c)))
		return rc;
	CK(ntc::launch_gen_packed(seed, first, n, L, mode, U, stride_words, d_words, c->stream));
	c->n_launches++;
	return NTC_OK;
}

int ntc_stats(ntc_ctx* c, uint64_t* n_launches, uint64_t* n_batches)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	if (n_launches) *n_launches = c->n_launches;
	if (n_batches) *n_batches = c->n_batches;
	return NTC_OK;
}

int ntc_stage_times(ntc_ctx* c, double* ms3)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c || !ms3)
		return set_err(NTC_EINVAL, "ntc_stage_times: bad argument");
	for (int i = 0; i < 3; i++) {
		ms3[i] = c->stage_ms[i];
		c->stage_ms[i] = 0;
	}
	return NTC_OK;
}

int ntc_kernel_time(ntc_ctx* c, double* ms_total, uint64_t* n_timed)
{
	std::unique_lock<std::recursive_mutex> lk_;
	if (c)
		lk_ = std::unique_lock<std::recursive_mutex>(c->mu);
	if (!c)
		return set_err(NTC_EINVAL, "null context");
	if (ms_total) *ms_total = c->kernel_ms;
	if (n_timed) *n_timed = c->n_timed;
	c->kernel_ms = 0;
	c->n_timed = 0;
	return NTC_OK;
}

} // extern "C"
