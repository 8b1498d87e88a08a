// sketch_kernels.cu -- general sm_100a kernels of the ntCard sketch path:
//   * piece bookkeeping for ragged / long records,
//   * roll64: one thread per piece, 64-bit NTF64/NTR64 recurrence in registers (any k, any s),
//   * finish: narrow uint32 counters mod 2^16 + counter-value histogram,
//   * the deterministic synthetic read generator (bench / tests).
// The bit-sliced fast kernel lives in bitslice_kernel.cu.
#include <cub/device/device_scan.cuh>

#include "launch.h"
#include "sketch_common.cuh"

namespace ntc {

// ------------------------------------------------------------------------------------------------
// piece bookkeeping
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t rec_offset(const uint32_t* __restrict__ off, uint32_t stride, uint32_t rec)
{
	return off ? (uint64_t)__ldg(off + rec) : (uint64_t)rec * stride;
}

// The record's length word, clamped to what its words can hold: a corrupt length (larger than 16 x (record words - 1)) must not
// send a kernel past the record -- the scan kernel clamps the same way (scan_kernel.cuh), so all paths agree on malformed input.
__device__ __forceinline__ uint32_t rec_len(const uint32_t* __restrict__ words, const uint32_t* __restrict__ off, uint32_t stride, uint32_t rec)
{
	const uint64_t o = rec_offset(off, stride, rec);
	const uint64_t cap = off ? (uint64_t)__ldg(off + rec + 1) - o : stride;
	const uint64_t room = cap ? (cap - 1) * 16 : 0;
	const uint32_t len = __ldg(words + o);
	return (uint64_t)len > room ? (uint32_t)room : len;
}

__global__ void piece_count_kernel(const uint32_t* __restrict__ words, const uint32_t* __restrict__ off, uint32_t stride,
    uint32_t n_rec, uint32_t kmin, uint32_t* __restrict__ n_pieces /* [n_rec + 1] */)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n_rec)
		return;
	uint32_t np = 0;
	if (i < n_rec) {
		uint32_t len = rec_len(words, off, stride, i);
		if (len >= kmin)
			np = (len - kmin + 1 + PIECE_STARTS - 1) / PIECE_STARTS;
	}
	n_pieces[i] = np; // entry n_rec = 0 so the exclusive scan leaves the total there
}

__global__ void piece_fill_kernel(const uint32_t* __restrict__ piece_first, uint32_t n_rec, uint32_t* __restrict__ piece_rec)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rec)
		return;
	uint32_t a = piece_first[i], b = piece_first[i + 1];
	for (uint32_t p = a; p < b; p++)
		piece_rec[p] = i;
}

// ------------------------------------------------------------------------------------------------
// re-tiling of long ragged records (long reads, N-split contigs) into fixed-length pieces, so that they can take
// the scan -> hit -> apply pipeline, which wants uniform-stride batches of equal-length records.
// Piece j of a record covers bases [j*D, j*D + Lp): Lp = 16 * (stride - 1) bases, D = Lp - kmax + 1 k-mer starts; the
// pipeline hashes the k-mers starting in the first D positions of every piece (for every k <= kmax that keeps each window
// inside its piece and counts each start once).  What is left of a record behind its last full piece -- bases
// [n_full * D, len), if at least kmin long -- becomes a "tail" record of a ragged batch for the general kernel.
// ------------------------------------------------------------------------------------------------
__global__ void retile_count_kernel(const uint32_t* __restrict__ words, const uint32_t* __restrict__ off, uint32_t n_rec, uint32_t Lp, uint32_t D,
    uint32_t kmin, uint32_t* __restrict__ n_full /* [n_rec + 1] */, uint32_t* __restrict__ has_tail, uint32_t* __restrict__ tail_words)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n_rec)
		return;
	uint32_t nf = 0, ht = 0, tw = 0;
	if (i < n_rec) {
		const uint32_t len = rec_len(words, off, 0, i);
		nf = len >= Lp ? (len - Lp) / D + 1 : 0;
		const uint32_t tl = len - nf * D;
		if (tl >= kmin && len >= kmin) {
			ht = 1;
			tw = 1 + (tl + 15) / 16;
		}
	}
	n_full[i] = nf;
	has_tail[i] = ht;
	tail_words[i] = tw;
}

// 16 bases starting at base `bo` of a record whose base words are b[0 .. nw)
__device__ __forceinline__ uint32_t bases16(const uint32_t* __restrict__ b, uint32_t nw, uint32_t bo)
{
	const uint32_t s = bo >> 4, sh = (bo & 15u) * 2u;
	const uint32_t lo = s < nw ? __ldg(b + s) : 0u, hi = s + 1 < nw ? __ldg(b + s + 1) : 0u;
	return __funnelshift_r(lo, hi, sh);
}

// one warp per record: the lanes write the record's output words
__global__ void __launch_bounds__(256) retile_fill_kernel(const uint32_t* __restrict__ words, const uint32_t* __restrict__ off, uint32_t n_rec,
    uint32_t Lp, uint32_t D, uint32_t stride, const uint32_t* __restrict__ full_first, const uint32_t* __restrict__ tail_idx,
    const uint32_t* __restrict__ tail_first, uint32_t* __restrict__ out_uniform, uint32_t* __restrict__ out_tail_words,
    uint32_t* __restrict__ out_tail_off)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t i = warp; i < n_rec; i += nwarps) {
		const uint32_t* r = words + __ldg(off + i);
		const uint32_t len = rec_len(words, off, 0, i), nw = (len + 15) / 16;
		const uint32_t f0 = full_first[i], nf = full_first[i + 1] - f0;
		uint32_t* ou = out_uniform + (size_t)f0 * stride;
		for (uint32_t x = lane; x < nf * stride; x += 32) {
			const uint32_t j = x / stride, w = x - j * stride;
			ou[x] = w == 0 ? Lp : bases16(r + 1, nw, j * D + (w - 1) * 16);
		}
		if (tail_idx[i + 1] != tail_idx[i]) {
			const uint32_t t0 = tail_first[i], tw = tail_first[i + 1] - t0, tl = len - nf * D;
			if (lane == 0)
				out_tail_off[tail_idx[i]] = t0;
			for (uint32_t w = lane; w < tw; w += 32) {
				uint32_t v = tl;
				if (w) {
					v = bases16(r + 1, nw, nf * D + (w - 1) * 16);
					const uint32_t left = tl - (w - 1) * 16;
					if (left < 16)
						v &= (1u << (2 * left)) - 1u;
				}
				out_tail_words[t0 + w] = v;
			}
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		out_tail_off[tail_idx[n_rec]] = tail_first[n_rec]; // end of the last tail record
}

// ------------------------------------------------------------------------------------------------
// roll64: the straightforward kernel.  ntRead's loops (ntcard.cpp:147-158) with the iterator's
// rolling update (ntHashIterator.hpp:85 -> NTC64, nthash.hpp:275-279); records hold valid bases
// only, so the N-restart branch (ntHashIterator.hpp:80-83) never fires on the device.
// ------------------------------------------------------------------------------------------------
template <bool kRecordIsPiece>
__global__ void __launch_bounds__(256) roll64_kernel(const uint32_t* __restrict__ words, const uint32_t* __restrict__ off,
    uint32_t stride, uint32_t n_rec, const uint32_t* __restrict__ piece_first, const uint32_t* __restrict__ piece_rec,
    const DevParams* __restrict__ P, uint32_t* __restrict__ counters, unsigned long long* __restrict__ f1, uint32_t kmask)
{
	__shared__ DevParams sp;
	__shared__ unsigned long long s_f1[NTC_MAX_K];
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(P);
		uint32_t* dst = reinterpret_cast<uint32_t*>(&sp);
		for (uint32_t i = threadIdx.x; i < sizeof(DevParams) / 4; i += blockDim.x)
			dst[i] = src[i];
		if (threadIdx.x < NTC_MAX_K)
			s_f1[threadIdx.x] = 0;
	}
	__syncthreads();
	const uint32_t n_pieces = kRecordIsPiece ? n_rec : piece_first[n_rec];
	const uint64_t tab_stride = (uint64_t)2 << sp.rBits;
	for (uint32_t piece = blockIdx.x * blockDim.x + threadIdx.x; piece < n_pieces; piece += gridDim.x * blockDim.x) {
		uint32_t rec, a;
		if (kRecordIsPiece) {
			rec = piece;
			a = 0;
		} else {
			rec = piece_rec[piece];
			a = (piece - piece_first[rec]) * PIECE_STARTS;
		}
		const uint32_t* r = words + rec_offset(off, stride, rec);
		const uint32_t len = rec_len(words, off, stride, rec);
		for (uint32_t ki = 0; ki < sp.nK; ki++) {
			if (!((kmask >> ki) & 1u))
				continue;
			uint32_t n = sp.gap ? process_piece_gap(r + 1, len, a, sp.k[ki], sp.gap, sp.gap_a31, sp.gap_a33, sp.tab[ki], sp.gtab,
			                          counters + ki * tab_stride, sp.rBits, sp.sBits)
			                    : process_piece_k(r + 1, len, a, sp.k[ki], sp.tab[ki], counters + ki * tab_stride, sp.rBits, sp.sBits);
			if (n)
				atomicAdd(&s_f1[ki], (unsigned long long)n);
		}
	}
	__syncthreads();
	if (threadIdx.x < sp.nK && s_f1[threadIdx.x])
		atomicAdd(f1 + threadIdx.x, s_f1[threadIdx.x]); // totKmer, ntcard.cpp:155,466
}

// ------------------------------------------------------------------------------------------------
// finish: narrow + counter-value histogram (compEst's first loop, ntcard.cpp:245-247)
// ------------------------------------------------------------------------------------------------
constexpr uint32_t HIST_SMEM_BINS = 2048;

// One CTA handles `chunk` consecutive counters of one table.  p_hist[table][v] for v >= 1 only; the
// host derives p[0] = 2^rBits - sum(v >= 1).
__global__ void __launch_bounds__(256) narrow_hist_kernel(const uint32_t* __restrict__ counters, uint64_t n_per_table,
    uint32_t chunk, uint16_t* __restrict__ narrow /* may be null */, uint32_t* __restrict__ p_hist /* may be null */)
{
	__shared__ uint32_t sh[HIST_SMEM_BINS];
	for (uint32_t i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
		sh[i] = 0;
	__syncthreads();
	const uint64_t chunks_per_table = (n_per_table + chunk - 1) / chunk;
	const uint64_t table = blockIdx.x / chunks_per_table;
	const uint64_t c0 = (blockIdx.x % chunks_per_table) * chunk;
	const uint64_t c1 = min(c0 + chunk, n_per_table);
	const uint32_t* src = counters + table * n_per_table;
	uint32_t* ph = p_hist ? p_hist + table * 65536 : nullptr;
	for (uint64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
		const uint32_t v = src[i] & 0xFFFFu; // uint16_t wrap of ntcard.cpp:143
		if (narrow)
			narrow[table * n_per_table + i] = (uint16_t)v;
		if (ph && v) {
			if (v < HIST_SMEM_BINS)
				atomicAdd(&sh[v], 1u);
			else
				atomicAdd(&ph[v], 1u);
		}
	}
	__syncthreads();
	if (ph)
		for (uint32_t i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
			if (sh[i])
				atomicAdd(&ph[i], sh[i]);
}

// Histogram of an arbitrary chunk-aligned range [first, first+n) of the flat counter array (the slice a
// rank owns after a reduce-scatter).  src points at element `first`.
__global__ void __launch_bounds__(256) hist_range_kernel(const uint32_t* __restrict__ src, uint64_t first, uint64_t n, uint32_t rBits,
    uint32_t chunk, uint32_t* __restrict__ p_hist)
{
	__shared__ uint32_t sh[HIST_SMEM_BINS];
	for (uint32_t i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
		sh[i] = 0;
	__syncthreads();
	const uint64_t c0 = (uint64_t)blockIdx.x * chunk;
	const uint64_t c1 = min(c0 + chunk, n);
	uint32_t* ph = p_hist + ((first + c0) >> rBits) * 65536; // the chunk lies inside one table
	for (uint64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
		const uint32_t v = src[i] & 0xFFFFu;
		if (v) {
			if (v < HIST_SMEM_BINS)
				atomicAdd(&sh[v], 1u);
			else
				atomicAdd(&ph[v], 1u);
		}
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
		if (sh[i])
			atomicAdd(&ph[i], sh[i]);
}

// ------------------------------------------------------------------------------------------------
// synthetic reads: word w (32 bases) of read i under seed S is mix64((S<<48)^(i<<12)^w),
// base j of the word = (v >> 2j) & 3 -- which already IS the 2-bit packing.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ULL;
	uint64_t z = x;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t gen_base(uint64_t S, uint64_t src, uint32_t j)
{
	return (uint32_t)(mix64((S << 48) ^ (src << 12) ^ (uint64_t)(j >> 5)) >> (2 * (j & 31))) & 3u;
}

__global__ void gen_packed_kernel(uint64_t S, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U, uint32_t stride,
    uint32_t* __restrict__ words)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t rec = t / stride;
	const uint32_t w = (uint32_t)(t % stride);
	if (rec >= n)
		return;
	uint32_t out = 0;
	if (w == 0) {
		out = L;
	} else {
		const uint32_t j0 = (w - 1) * 16;
		if (j0 < L) {
			const uint64_t i = first + rec;
			const bool rep = (mode == 1 && U != 0);
			const uint64_t src = rep ? i % U : i;
			const bool rc = rep && ((i / U) & 1);
			const uint32_t nb = min(16u, L - j0);
			if (!rc) {
				uint64_t v = mix64((S << 48) ^ (src << 12) ^ (uint64_t)(j0 >> 5));
				out = (uint32_t)(v >> (2 * (j0 & 31)));
				if (nb < 16)
					out &= (1u << (2 * nb)) - 1;
			} else {
				for (uint32_t j = 0; j < nb; j++)
					out |= (3u - gen_base(S, src, L - 1 - (j0 + j))) << (2 * j);
			}
		}
	}
	words[rec * stride + w] = out;
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
static inline unsigned grid_for(uint64_t n, unsigned block, unsigned cap)
{
	uint64_t g = (n + block - 1) / block;
	if (g < 1) g = 1;
	if (g > cap) g = cap;
	return (unsigned)g;
}

size_t piece_scan_temp_bytes(uint32_t n_items)
{
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_items);
	return bytes;
}

cudaError_t launch_piece_tables(const BatchView& b, uint32_t kmin, uint32_t* d_piece_first, uint32_t* d_piece_rec, void* d_tmp,
    size_t tmp_bytes, cudaStream_t st)
{
	const uint32_t n = b.n_rec + 1;
	piece_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(b.words, b.off, b.stride, b.n_rec, kmin, d_piece_first);
	cudaError_t e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_piece_first, d_piece_first, (int)n, st);
	if (e != cudaSuccess)
		return e;
	piece_fill_kernel<<<(b.n_rec + 255) / 256, 256, 0, st>>>(d_piece_first, b.n_rec, d_piece_rec);
	return cudaGetLastError();
}

// uniform-stride batch whose stride is not a multiple of 4 words (tightly packed: 44 bytes per 150 bp read) -> the same
// records at the next multiple of 4, which is what the scan kernel's 128-bit loads need
__global__ void __launch_bounds__(256) restride_kernel(const uint32_t* __restrict__ in, uint32_t stride_in, uint32_t stride_out, uint64_t n_out_words,
    uint32_t* __restrict__ out)
{
	for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n_out_words; x += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t rec = x / stride_out;
		const uint32_t w = (uint32_t)(x - rec * stride_out);
		out[x] = w < stride_in ? __ldg(in + rec * stride_in + w) : 0u;
	}
}

cudaError_t launch_restride(const uint32_t* d_in, uint32_t stride_in, uint32_t stride_out, uint32_t n_rec, uint32_t* d_out, int n_sm, cudaStream_t st)
{
	const uint64_t n = (uint64_t)n_rec * stride_out;
	if (n == 0)
		return cudaSuccess;
	restride_kernel<<<grid_for(n, 256, (unsigned)n_sm * 32u), 256, 0, st>>>(d_in, stride_in, stride_out, n, d_out);
	return cudaGetLastError();
}

// headerless uniform batch (ntc_submit_bases: n records of wpr base words each, all `len` bases long, NO length words: 40 instead of 44
// bytes per 150 bp read over PCIe) -> length-prefixed records at a stride that is a multiple of 4 words
__global__ void __launch_bounds__(256) add_headers_kernel(const uint32_t* __restrict__ in, uint32_t wpr, uint32_t len, uint32_t stride_out,
    uint64_t n_out_words, uint32_t* __restrict__ out)
{
	for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n_out_words; x += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t rec = x / stride_out;
		const uint32_t w = (uint32_t)(x - rec * stride_out);
		out[x] = w == 0 ? len : w <= wpr ? __ldg(in + rec * wpr + (w - 1)) : 0u;
	}
}

cudaError_t launch_add_headers(const uint32_t* d_in, uint32_t wpr, uint32_t len, uint32_t stride_out, uint32_t n_rec, uint32_t* d_out, int n_sm,
    cudaStream_t st)
{
	const uint64_t n = (uint64_t)n_rec * stride_out;
	if (n == 0)
		return cudaSuccess;
	add_headers_kernel<<<grid_for(n, 256, (unsigned)n_sm * 32u), 256, 0, st>>>(d_in, wpr, len, stride_out, n, d_out);
	return cudaGetLastError();
}

// ragged batch of short records -> the same records zero-padded to one stride (a multiple of 4 words), so that they can take the
// pipeline as tiles of mixed lengths (scan_kernel.cuh)
__global__ void __launch_bounds__(256) pad_ragged_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ off, uint32_t stride_out,
    uint64_t n_out_words, uint32_t* __restrict__ out)
{
	for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n_out_words; x += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t rec = (uint32_t)(x / stride_out), w = (uint32_t)(x - (uint64_t)rec * stride_out);
		const uint32_t o = __ldg(off + rec), nw = __ldg(off + rec + 1) - o;
		out[x] = w < nw ? __ldg(in + o + w) : 0u;
	}
}

cudaError_t launch_pad_ragged(const uint32_t* d_in, const uint32_t* d_off, uint32_t n_rec, uint32_t stride_out, uint32_t* d_out, int n_sm,
    cudaStream_t st)
{
	const uint64_t n = (uint64_t)n_rec * stride_out;
	if (n == 0)
		return cudaSuccess;
	pad_ragged_kernel<<<grid_for(n, 256, (unsigned)n_sm * 32u), 256, 0, st>>>(d_in, d_off, stride_out, n, d_out);
	return cudaGetLastError();
}

// ntc_check_offsets on the device, for ragged batches that are already resident there: out[0] = words of the longest record,
// out[1] != 0 when the offsets are not a non-decreasing sequence from 0 to at most n_words (out zeroed by the caller)
__global__ void __launch_bounds__(256) check_offsets_kernel(const uint32_t* __restrict__ off, uint32_t n_rec, uint64_t n_words, uint32_t* __restrict__ out)
{
	uint32_t mx = 0, bad = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rec; i += gridDim.x * blockDim.x) {
		const uint32_t a = __ldg(off + i), b = __ldg(off + i + 1);
		if (b < a || (uint64_t)b > n_words || (i == 0 && a != 0))
			bad = 1;
		else
			mx = max(mx, b - a);
	}
	mx = __reduce_max_sync(0xFFFFFFFFu, mx);
	bad = __reduce_max_sync(0xFFFFFFFFu, bad);
	if ((threadIdx.x & 31u) == 0) {
		if (mx)
			atomicMax(out, mx);
		if (bad)
			atomicMax(out + 1, 1u);
	}
}

cudaError_t launch_check_offsets(const uint32_t* d_off, uint32_t n_rec, uint64_t n_words, uint32_t* d_out, int n_sm, cudaStream_t st)
{
	check_offsets_kernel<<<grid_for(n_rec, 256, (unsigned)n_sm * 8u), 256, 0, st>>>(d_off, n_rec, n_words, d_out);
	return cudaGetLastError();
}

__global__ void stride_offsets_kernel(uint32_t stride, uint32_t n_rec, uint32_t* __restrict__ off /* [n_rec + 1] */)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i <= n_rec)
		off[i] = i * stride;
}

// word offsets of a uniform-stride batch, so that it can go through the kernels written for ragged batches
cudaError_t launch_stride_offsets(uint32_t stride, uint32_t n_rec, uint32_t* d_off, cudaStream_t st)
{
	stride_offsets_kernel<<<(n_rec + 1 + 255) / 256, 256, 0, st>>>(stride, n_rec, d_off);
	return cudaGetLastError();
}

cudaError_t launch_retile_count(const BatchView& b, uint32_t Lp, uint32_t D, uint32_t kmin, uint32_t* d_n_full, uint32_t* d_has_tail,
    uint32_t* d_tail_words, void* d_tmp, size_t tmp_bytes, cudaStream_t st)
{
	const uint32_t n = b.n_rec + 1;
	retile_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(b.words, b.off, b.n_rec, Lp, D, kmin, d_n_full, d_has_tail, d_tail_words);
	cudaError_t e;
	if ((e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_n_full, d_n_full, (int)n, st)) != cudaSuccess ||
	    (e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_has_tail, d_has_tail, (int)n, st)) != cudaSuccess ||
	    (e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_tail_words, d_tail_words, (int)n, st)) != cudaSuccess)
		return e;
	return cudaGetLastError();
}

cudaError_t launch_retile_fill(const BatchView& b, uint32_t Lp, uint32_t D, uint32_t stride, const uint32_t* d_full_first,
    const uint32_t* d_tail_idx, const uint32_t* d_tail_first, uint32_t* d_out_uniform, uint32_t* d_out_tail_words, uint32_t* d_out_tail_off,
    int n_sm, cudaStream_t st)
{
	const unsigned cap = (unsigned)n_sm * 16u;
	const unsigned grid = grid_for((uint64_t)b.n_rec * 32, 256, cap);
	retile_fill_kernel<<<grid, 256, 0, st>>>(b.words, b.off, b.n_rec, Lp, D, stride, d_full_first, d_tail_idx, d_tail_first, d_out_uniform,
	    d_out_tail_words, d_out_tail_off);
	return cudaGetLastError();
}

cudaError_t launch_roll64(const BatchView& b, bool record_is_piece, uint64_t n_pieces_bound, const uint32_t* d_piece_first,
    const uint32_t* d_piece_rec, const DevParams* d_params, uint32_t* d_counters, unsigned long long* d_f1, uint32_t kmask,
    int n_sm, cudaStream_t st)
{
	// enough CTAs for every SM to hold its resident set, a multiple of the SM count
	const unsigned cap = (unsigned)n_sm * 8u * 16u;
	unsigned grid = grid_for(record_is_piece ? b.n_rec : n_pieces_bound, 256, cap);
	if (record_is_piece)
		roll64_kernel<true><<<grid, 256, 0, st>>>(b.words, b.off, b.stride, b.n_rec, nullptr, nullptr, d_params, d_counters, d_f1, kmask);
	else
		roll64_kernel<false><<<grid, 256, 0, st>>>(b.words, b.off, b.stride, b.n_rec, d_piece_first, d_piece_rec, d_params,
		    d_counters, d_f1, kmask);
	return cudaGetLastError();
}

cudaError_t launch_narrow_hist(const uint32_t* d_counters, uint32_t n_tables, uint64_t n_per_table, uint16_t* d_narrow,
    uint32_t* d_phist, cudaStream_t st)
{
	const uint32_t chunk = (uint32_t)(n_per_table < 65536 ? n_per_table : 65536);
	const uint64_t chunks_per_table = (n_per_table + chunk - 1) / chunk;
	narrow_hist_kernel<<<(unsigned)(n_tables * chunks_per_table), 256, 0, st>>>(d_counters, n_per_table, chunk, d_narrow, d_phist);
	return cudaGetLastError();
}

cudaError_t launch_hist_range(const uint32_t* d_src, uint64_t first, uint64_t n, uint32_t rBits, uint32_t* d_phist, cudaStream_t st)
{
	const uint64_t per_table = (uint64_t)1 << rBits;
	const uint32_t chunk = (uint32_t)(per_table < 65536 ? per_table : 65536);
	if (n == 0)
		return cudaSuccess;
	if (first % chunk || n % chunk)
		return cudaErrorInvalidValue;
	hist_range_kernel<<<(unsigned)(n / chunk), 256, 0, st>>>(d_src, first, n, rBits, chunk, d_phist);
	return cudaGetLastError();
}

cudaError_t launch_gen_packed(uint64_t S, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U, uint32_t stride,
    uint32_t* d_words, cudaStream_t st)
{
	const uint64_t total = n * stride;
	if (total == 0)
		return cudaSuccess;
	gen_packed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S, first, n, L, mode, U, stride, d_words);
	return cudaGetLastError();
}

} // namespace ntc
