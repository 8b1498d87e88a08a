// bitslice_launch.h -- host-side launch interface of the bit-sliced kernel (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ntc {
struct DevParams;
namespace bs {

struct BsLaunch {          // per-k launch constants, passed by value
	uint32_t k, ki, rBits;
	uint32_t ring;                  // positions in a scan warp's shared-memory plane ring (power of two >= k + 16)
	uint32_t nbuf;                  // mask buffers (hand-off units in flight) per scan warp
	uint32_t pairs;                 // active scan warps per CTA (each with two hit warps)
	uint32_t dbg;                   // experiments: bit0 = hit warps skip all work, bit1 = no RED
	uint32_t F0[31], R0[31];        // initial bit-sliced state (bitslice_core.cuh init_state)
	uint64_t rot_a, rot_b;          // byte m: (k%32 + 32m) % 31 and % 33 for block m of the hit path (k < 288)
};

struct BsArgs {
	const uint32_t* words;
	uint32_t stride, n_rec;
	BsLaunch L;
	const uint4* d_tab;             // [8][256] byte tables of the hit path
	const DevParams* d_params;
	uint32_t* ctr_k;                // counters of this k: [2][2^rBits]
	unsigned long long* f1_k;
	unsigned grid, pairs;           // CTAs; (scan warp, hit warp) pairs per CTA
	size_t smem_bytes;
	cudaStream_t stream;
};

constexpr size_t kTabBytes = 8 * 256 * 16;
constexpr size_t kMaskBytes = 16 * 128;               // one hand-off unit's masks: [16 positions][32 lanes] words
constexpr int kHitWarpsPerScan = 3;                   // must equal bs::kHitWarps in bitslice_kernel.cuh
constexpr int kQueueCap = 288;                        // hit queue entries per hit warp: one HALF body (16 positions), expected 256 at s=7
constexpr size_t kSmemMax = 232448;                   // 227 KB opt-in limit per CTA on sm_100

// True when a kernel for (k mod 31, sBits) was compiled in.
bool have_kernel(unsigned k, unsigned sBits);
// Launch for one k.  Returns cudaErrorInvalidValue when no instantiation exists.
cudaError_t launch(unsigned k, unsigned sBits, const BsArgs& a);
// Fill the hit-path byte tables (host): tab[j*256+b] = {FB lo, FB hi, RB lo, RB hi}.
void build_tables(uint32_t* tab /* 8*256*4 words */);

} // namespace bs
} // namespace ntc
