// pipeline.h -- host-side interface of the three-stage sketch pipeline (internal):
//
//   S  scan_kernel<KM,S>   bit-sliced upper-ring filter over a uniform-stride batch -> one 32-slot mask word per
//                          (tile, k-mer position, lane) in HBM, the candidate count, per-tile info
//   H  hit_kernel          candidates -> full 64-bit canonical hash -> ntComp -> counter index, appended to a
//                          BINNED hit log (pool of 1 KB blocks, one list of blocks per 32 MiB slice of the sketch)
//   A  apply_kernel        per slice: bring the slice into L2 (write zeros on the first flush after a reset, read
//                          it otherwise), then RED.ADD the slice's log entries -- L2-resident atomics run ~8x
//                          faster than random atomics into the 1 GiB sketch in HBM
//   F  fallback_kernel     tiles the scan kernel flagged (records of different lengths AND the all-A k-mer of the padding is sampled
//                          for this k): 64-bit recurrence, direct RED
//
// Exactness never depends on capacities: when the pool is exhausted H increments the sketch directly, and the
// conditional flush in front of H (apply_kernel with force = 0) guarantees the sketch is materialised (zeroed)
// whenever that can happen.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ntc {
struct DevParams;
namespace pl {

constexpr uint32_t kBlkEntries = 256;         // log entries per pool block (1 KB)
constexpr uint32_t kWireBlkWords = 260;        // a log block on the wire (multi-GPU exchange): 256 entries, fill, 3 words of padding
constexpr uint32_t kBlkSlack = 96;            // a block with fewer free entries than this is closed at the end of a round
constexpr uint32_t kVoid = 0xFFFFFFFFu;
constexpr uint32_t kTileFlag = 0xFFFFFFFFu;   // tile_info value: not uniform -> fallback kernel
constexpr uint32_t kTileDefer = 0x80000000u;  // fused kernel: tile_info = kTileDefer | first k-mer start still to be hashed (pool ran out
                                              // while the sketch was not materialised: second pass after the conditional flush)
constexpr uint32_t kFusedMaxPos = 2048;       // fused kernel: k-mer starts per record (11 bits of a 16-bit candidate word)
constexpr uint32_t kFusedWarps = 8;
// words per saved state of a hit group / fused-kernel warp: [epoch][flush count + 1][5 x nbins]
__host__ __device__ constexpr size_t gstate_row(uint32_t nbins) { return 2 + 5 * (size_t)nbins; }
constexpr uint32_t kTileRecs = 1024;
constexpr uint32_t kMaxBins = 64;             // slices per k
constexpr uint32_t kQueueCap = 4096;          // candidates per round of a hit group
constexpr uint32_t kGroupThreads = 256;
constexpr uint32_t kApplyThreads = 512;
constexpr size_t kSmemMax = 232448;           // 227 KB opt-in limit per CTA on sm_100

// control words (device)
// NFLAG and NDEFER are adjacent: one 8-byte memset clears both at the start of a batch
enum { CTL_NEXT = 0, CTL_STATE = 1, CTL_NFLAG = 2, CTL_NDEFER = 3, CTL_TICKET = 4, CTL_FLUSHES = 5, CTL_DIRECT = 6, CTL_WORDS = 8 };

struct Pool {                    // device pointers + geometry, passed by value
	uint32_t* entries;           // [n_blocks][kBlkEntries] counter indices inside one k's [2][2^rBits] sub-sketch
	uint32_t* slice_blocks;      // [n_slices][slice_cap] block id << 9 | entries in the block
	uint32_t* slice_nblk;        // [n_slices]
	uint32_t* zero_done;         // [n_slices] CTAs that zeroed their share of the slice (apply kernel)
	uint32_t* apply_done;        // [n_slices] CTAs that applied their share of the slice's log
	uint32_t* ctl;               // [CTL_WORDS]
	unsigned long long* cand;    // candidate k-mers of the batch being processed (upper bound of its log entries)
	uint32_t n_blocks, slice_cap, n_slices, nbins, bin_shift, rBits, nK;
	uint32_t* gstate;            // [nK][max_groups][gstate_row(nbins)]: open / spare blocks of every hit group, kept over launches
	uint32_t max_groups, epoch;  // epoch: bumped by ntc_reset (host); with the full 32-bit CTL_FLUSHES it dates gstate
	uint32_t ahead;              // apply kernel: slices the stagers may run ahead of the appliers (L2 footprint)
};

constexpr uint32_t kMaxPeers = 8;
struct PeerLogs {                // the hit logs of all ranks of one node, mapped through CUDA IPC (this rank's own included)
	const uint32_t* entries[kMaxPeers];
	const uint32_t* slice_blocks[kMaxPeers];
	const uint32_t* slice_nblk[kMaxPeers];
	uint32_t n, self;            // number of ranks, this rank
};

struct HashK {                   // per-k constants of the block-wise full hash (hit_hash.cuh: make_hashk)
	uint32_t k, tprime, nblk;    // k = tprime + 32 (nblk - 1), 1 <= tprime <= 32
	uint32_t head_ra, head_rb;   // srol amounts (mod 31, mod 33) equal to sror^(32 - tprime)
	uint64_t head_c, head_d;     // what the 32 - tprime masked-out codes of the head block contribute
	uint64_t rot_a, rot_b;       // byte m-1: (tprime + 32 (m-1)) % 31 and % 33
};

struct ScanLaunch {              // per-k constants of the scan kernel, passed by value
	uint32_t k, ring, nwarps;    // ring: positions in a warp's plane ring (k + 16 rounded up to a multiple of 16)
	uint32_t npos_max;           // mask rows per tile
	uint32_t start_limit;        // != 0: hash only the k-mers starting in the first start_limit positions of a record (re-tiled pieces)
	uint32_t prefetch;           // cp.async.bulk.prefetch.L2 of the warp's next tile: 0 = none, 1 = half way through the current one, 2 = one column before its end
	uint32_t mixed_ok;           // != 0: tiles of records of different lengths may be scanned at the longest length (the all-A k-mer is not sampled)
	uint32_t F0[31], R0[31];     // initial bit-sliced state (bitslice_core.cuh init_state)
};

struct ScanArgs {
	const uint32_t* words;
	uint32_t stride, n_rec;
	ScanLaunch L;
	uint32_t* masks;             // [n_tiles][npos_max][32]
	uint32_t* tile_info;         // [n_tiles]: k-mer positions per record (uniform tile), 0, or kTileFlag
	unsigned long long* f1_k;
	unsigned long long* cand;
	uint32_t* ctl;
	const uint32_t* hll_min = nullptr; // nthll pre-filter: device word holding the smallest register (launch_hll_min)
	unsigned grid;
	size_t smem_bytes;
	cudaStream_t stream;
};

struct FusedArgs {               // the fused sketch kernel (fused_kernel.cuh): scan + full hash + log append in one launch
	const uint32_t* words;
	uint32_t stride, n_rec, n_tiles;
	ScanLaunch L;                // ring here also sizes the per-warp index buffer that aliases the planes
	uint32_t ki, sBits, pass;    // pass 0: every tile (counts F1); pass 1: only the tiles pass 0 deferred
	uint32_t qlane;              // candidate slots per lane of a warp's queue (<= 2 * (ring + 3))
	uint32_t dbg;                // timing experiments only (NTC_FUSED_DBG): 1 = skip the hash + append phase, 2 = skip the emission (results are wrong)
	const uint4* d_tab;          // [8][256] byte tables of the full hash
	HashK hk;
	uint32_t* ctr_k;             // counters of this k (direct increments once the sketch is materialised and the pool is full)
	Pool pool;
	uint32_t* tile_info;         // [n_tiles]: 0 done, kTileFlag -> fallback kernel, kTileDefer | p -> second pass
	unsigned long long* f1_k;
	unsigned grid;
	size_t smem_bytes;
	cudaStream_t stream;
};

struct HitArgs {
	const uint32_t* words;
	uint32_t stride, n_rec, n_tiles;
	uint32_t k, ki, sBits, npos_max;
	uint32_t rows_per_unit, units_per_tile, tiles_per_unit, n_units; // units_per_tile > 1 xor tiles_per_unit > 1 (or both 1)
	const uint32_t* masks;
	const uint32_t* tile_info;   // k-mer positions per record of every tile (0 / kTileFlag: no rows to read)
	const uint4* d_tab;          // [8][256] byte tables of the full hash
	HashK hk;
	uint32_t* ctr_k;             // counters of this k
	Pool pool;
	uint32_t mask_prefetch = 0;  // != 0: a group asks for its next unit's mask rows (prefetch.global.L2) before it works on the current one
	cudaStream_t stream;
};

// True when a scan kernel for (k mod 31, sBits) was compiled in (Makefile BS_KMS; sBits 7 and 11).
bool have_scan_kernel(unsigned k, unsigned sBits);
// Fill the byte tables of the full hash (host): tab[j*256+b] = {FB lo, FB hi, RB lo, RB hi}.
void build_tables(uint32_t* tab /* 8*256*4 words */);
cudaError_t launch_scan(unsigned k, unsigned sBits, const ScanArgs& a);
cudaError_t launch_fused(unsigned k, unsigned sBits, const FusedArgs& a);
// nthll pre-filter (hll_kernels.cu): the scan kernel with "top T bits of the canonical hash are zero" as its predicate; T = 13, or
// T = 0: 5 / 7 / 9 / 11 / 13 bits by the smallest register as ScanArgs::hll_min holds it when the kernel starts
cudaError_t launch_hllscan(unsigned k, unsigned T, const ScanArgs& a);
size_t fused_smem_bytes(uint32_t ring, uint32_t nwarps, uint32_t qlane, uint32_t nbins);
bool hit_can_stage(uint32_t stride, uint32_t units_per_tile, uint32_t tiles_per_unit);
unsigned hit_groups_per_sm(bool staged);
cudaError_t launch_hit(const HitArgs& a, bool staged, unsigned ctas);
cudaError_t launch_batch_clear(const Pool& pool, cudaStream_t st); // candidate count and flagged-tile count of the batch := 0
cudaError_t launch_fallback(const uint32_t* words, uint32_t stride, uint32_t n_rec, uint32_t n_tiles, const uint32_t* tile_info,
    const DevParams* d_params, uint32_t ki, uint32_t* ctr_k, const uint32_t* ctl, int n_sm, cudaStream_t st);
// force = 0: flush only if the batch about to be hashed could overflow the pool while the sketch is still
// unmaterialised (or has flagged tiles); force = 1: flush whatever is pending and materialise the sketch.
// force = 2 (fused kernel): flush when tiles were deferred or flagged, or when a materialised sketch's pool is 3/4 full.
cudaError_t launch_apply(const Pool& pool, uint32_t* counters, int force, uint32_t reserve_blocks, unsigned grid, cudaStream_t st,
    const uint32_t* d_order = nullptr, uint32_t n_order = 0);
// Hit-log exchange between ranks: copy the blocks of runs of slices to contiguous buffers / register received blocks
// (already copied to pool blocks base_blk..) in the block lists of their slices.  runs: [n_runs][2] = {slice, first block}.
cudaError_t launch_export(const Pool& pool, const uint32_t* d_runs, uint32_t n_runs, uint32_t n_out, uint32_t* d_out, cudaStream_t st);
cudaError_t launch_import(const Pool& pool, const uint32_t* d_runs, uint32_t n_runs, uint32_t n_in, uint32_t base_blk, const uint32_t* d_in,
    cudaStream_t st);
cudaError_t launch_hist_slices(const Pool& pool, const uint32_t* counters, const uint32_t* d_order, uint32_t n_order, uint32_t* d_phist, cudaStream_t st);
int apply_max_grid(int n_sm);
// Multi-GPU reduction over peer memory: zero the owned slices and apply to them the log entries of ALL ranks (hit_kernels.cu).
cudaError_t launch_apply_owned(const Pool& pool, const PeerLogs& logs, uint32_t* counters, const uint32_t* d_order, uint32_t n_order,
    const long long* d_status, uint32_t status_idx, unsigned grid, cudaStream_t st);
cudaError_t launch_log_status(const Pool& pool, const unsigned long long* d_f1, uint32_t nK, int host_ok, long long* d_out, cudaStream_t st);

} // namespace pl
} // namespace ntc
