// launch.h -- host-callable launchers of the sm_100a kernels (internal; the public surface is
// include/ntcard_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ntc {

struct DevParams;

struct BatchView {          // a packed record stream resident in device memory
	const uint32_t* words;  // record i at words[off[i]] (or i*stride): [len][ceil(len/16) base words]
	const uint32_t* off;    // n_rec + 1 word offsets, or nullptr for uniform stride
	uint32_t stride;        // words per record when off == nullptr
	uint32_t n_rec;
	uint64_t n_words;
	uint32_t uniform_len;   // bases per record when the host knows all records are equal, else 0
	uint32_t max_rec_words; // ragged batches: words of the longest record when the host knows it, else 0
	uint32_t headerless_len = 0; // != 0: records carry no length word -- stride words of bases each, all this many bases long
};

size_t piece_scan_temp_bytes(uint32_t n_items);
cudaError_t launch_piece_tables(const BatchView& b, uint32_t kmin, uint32_t* d_piece_first, uint32_t* d_piece_rec, void* d_tmp,
    size_t tmp_bytes, cudaStream_t st);
// Long ragged records -> uniform pieces of Lp bases every D bases (+ ragged tails); see sketch_kernels.cu.  After
// launch_retile_count the three arrays hold exclusive prefix sums (entry n_rec = totals).
cudaError_t launch_retile_count(const BatchView& b, uint32_t Lp, uint32_t D, uint32_t kmin, uint32_t* d_n_full, uint32_t* d_has_tail,
    uint32_t* d_tail_words, void* d_tmp, size_t tmp_bytes, cudaStream_t st);
cudaError_t launch_retile_fill(const BatchView& b, uint32_t Lp, uint32_t D, uint32_t stride, const uint32_t* d_full_first,
    const uint32_t* d_tail_idx, const uint32_t* d_tail_first, uint32_t* d_out_uniform, uint32_t* d_out_tail_words, uint32_t* d_out_tail_off,
    int n_sm, cudaStream_t st);
cudaError_t launch_check_offsets(const uint32_t* d_off, uint32_t n_rec, uint64_t n_words, uint32_t* d_out /* [2], zeroed */, int n_sm, cudaStream_t st);
cudaError_t launch_stride_offsets(uint32_t stride, uint32_t n_rec, uint32_t* d_off, cudaStream_t st);
cudaError_t launch_restride(const uint32_t* d_in, uint32_t stride_in, uint32_t stride_out, uint32_t n_rec, uint32_t* d_out, int n_sm, cudaStream_t st);
cudaError_t launch_add_headers(const uint32_t* d_in, uint32_t wpr, uint32_t len, uint32_t stride_out, uint32_t n_rec, uint32_t* d_out, int n_sm,
    cudaStream_t st);
cudaError_t launch_pad_ragged(const uint32_t* d_in, const uint32_t* d_off, uint32_t n_rec, uint32_t stride_out, uint32_t* d_out, int n_sm,
    cudaStream_t st);
cudaError_t launch_roll64(const BatchView& b, bool record_is_piece, uint64_t n_pieces_bound, const uint32_t* d_piece_first,
    const uint32_t* d_piece_rec, const DevParams* d_params, uint32_t* d_counters, unsigned long long* d_f1, uint32_t kmask,
    int n_sm, cudaStream_t st);
// nthll's HyperLogLog registers (hll_kernels.cu): d_regs = 2^nBits one-byte registers (at least 4 bytes, 4-byte aligned)
cudaError_t launch_hll(const BatchView& b, bool record_is_piece, uint64_t n_pieces_bound, const uint32_t* d_piece_first,
    const uint32_t* d_piece_rec, const DevParams* d_params, uint32_t nBits, uint8_t* d_regs, unsigned long long* d_f1, int n_sm,
    cudaStream_t st);
// nthll fast path: mask words of the pre-filter scan -> full hash -> ntComp on the registers; smallest register
cudaError_t launch_hll_hit(const uint32_t* d_words, uint32_t stride, uint32_t n_rec, uint32_t n_tiles, uint32_t npos_max, const uint32_t* d_masks,
    const uint32_t* d_tile_info, const uint4* d_tab, uint32_t k, uint32_t nBits, uint8_t* d_regs, int n_sm, cudaStream_t st);
cudaError_t launch_hll_min(const uint8_t* d_regs, uint32_t nBits, uint32_t* d_out, cudaStream_t st);
cudaError_t launch_narrow_hist(const uint32_t* d_counters, uint32_t n_tables, uint64_t n_per_table, uint16_t* d_narrow,
    uint32_t* d_phist, cudaStream_t st);
cudaError_t launch_hist_range(const uint32_t* d_src, uint64_t first, uint64_t n, uint32_t rBits, uint32_t* d_phist, cudaStream_t st);
cudaError_t launch_gen_packed(uint64_t S, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U, uint32_t stride,
    uint32_t* d_words, cudaStream_t st);

} // namespace ntc
