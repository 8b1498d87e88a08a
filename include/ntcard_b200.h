/*
 * ntcard_b200.h -- C-ABI of the B200-native ntCard sketch path.
 *
 * The reference (bcgsc/ntCard v1.2.2) has no plugin / FFI layer; the seam this
 * library replaces is the FUNCTION boundary between its readers and its sketch:
 *
 *   ntRead(const string& seq, const vector<unsigned>& kList,
 *          uint16_t* t_Counter, size_t totKmer[])          ntcard.cpp:147-158
 *      called once per sequence by getEfq/getEfa/getEsm    ntcard.cpp:182,203,230
 *   ntComp(hVal, t_Counter)                                 ntcard.cpp:132-145
 *   t_Counter = new uint16_t[nK * nSamp * rBuck]()          ntcard.cpp:437-439
 *   totalKmers[k] += totKmer[k]                             ntcard.cpp:464-466
 *   compEst(const uint16_t* t_Counter, double& F0Mean,
 *           double fMean[])                                 ntcard.cpp:237-275
 *
 * A device call per read is meaningless, so the boundary is batch oriented:
 * the host parses reads, splits them at non-ACGTU characters (a window is hashed
 * iff all k chars are valid: ntHashIterator.hpp:66-67,80-83), 2-bit packs the
 * segments into a length-prefixed record stream, and submits batches.  The
 * device rolls canonical ntHash (NTC64, nthash.hpp:242-279) over every k-mer of
 * every record for every k, samples (ntComp) and increments the sketch.
 *
 * All entry points return 0 on success or a negative NTC_E* code;
 * ntc_last_error() gives the message (thread local).  There is NO CPU fallback:
 * every compute entry point fails with NTC_ENODEVICE when no CUDA device is
 * usable.  Plain pointers and sizes only -- no C++ or torch types.
 */
#ifndef NTCARD_B200_H
#define NTCARD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTC_OK 0
#define NTC_EINVAL (-1)    /* bad argument */
#define NTC_ENODEVICE (-2) /* no usable CUDA device / driver */
#define NTC_ECUDA (-3)     /* a CUDA call failed; see ntc_last_error() */
#define NTC_ENOMEM (-4)
#define NTC_ESTATE (-5)    /* call not valid in this state */

#define NTC_MAX_K 16       /* distinct k values per context */
#define NTC_NSAMP 2        /* opt::nSamp, ntcard.cpp:61 */

typedef struct ntc_ctx ntc_ctx;

/* ---- packed record stream -----------------------------------------------------
 * A batch is an array of 32-bit little-endian words.  Record i occupies words
 * [off[i], off[i+1]):
 *     word 0      : len  = number of bases (all valid: A C G T/U, either case)
 *     words 1..   : ceil(len/16) words, base j in bits [2(j%16), 2(j%16)+1] of
 *                   word 1 + j/16;  A=0 C=1 G=2 T/U=3 (complement = 3 - code)
 *     padding     : zero words up to off[i+1] (optional)
 * off == NULL means every record has the same word count `stride_words`
 * (off[i] = i * stride_words); stride_words % 4 == 0 enables the TMA path.
 * Algorithmic bytes per record = 4 + ceil(len/4)   (SURVEY.md 8d).            */

/* kernel selection for ntc_set_kernel(): */
#define NTC_KERNEL_AUTO 0
#define NTC_KERNEL_ROLL64 1   /* general: one thread per piece, 64-bit recurrence */
#define NTC_KERNEL_BITSLICE 2 /* 32 records bit-sliced per lane, upper-ring filter */

/* Create a context on CUDA device `device` for k values kList[0..nK) with a
 * sketch of NTC_NSAMP x 2^rBits counters per k, sampling on sBits leading bits
 * (opt::rBits / opt::sBits, ntcard.cpp:57-58; the caller applies the "sBits = 7
 * below 50 GB" rule of ntcard.cpp:430-431).  Replaces ntcard.cpp:437-439.
 * d_counters: optional caller-owned DEVICE buffer of nK*2*2^rBits uint32 (its content
 * is defined -- zeroed, then incremented -- from the first flush on; read it only after
 * ntc_counters_device / ntc_sync), e.g. a torch tensor that torch.distributed will all-reduce;
 * NULL lets the context allocate it.  cuda_stream: optional cudaStream_t on
 * which all device work is ordered (NULL = the context creates its own). */
int ntc_create(ntc_ctx** out, const unsigned* kList, unsigned nK, unsigned rBits, unsigned sBits, int device,
    void* d_counters, void* cuda_stream);
void ntc_destroy(ntc_ctx* ctx);

/* Zero the sketch and the k-mer totals (a fresh `new uint16_t[...]()`); the zeros are written
 * lazily, by the first flush. */
int ntc_reset(ntc_ctx* ctx);
int ntc_set_kernel(ntc_ctx* ctx, int kernel);
/* Gap seeds, the reference's -g N (stRead / stHashIterator / NTMSM64, ntcard.cpp:160-171, 407-413; nthash.hpp:620-678):
 * the middle `gap` bases of every k-mer are don't-cares, h = min(fh ^ Mf, rh ^ Mr).  One k only, gap and k of equal
 * parity, gap <= k - 2 (the reference's own checks, ntcard.cpp:382, 397).  gap = 0 switches back.  Batches submitted
 * afterwards use the general kernel. */
int ntc_set_gap(ntc_ctx* ctx, unsigned gap);

/* Submit one batch held in HOST memory (replaces n_rec calls of ntRead,
 * ntcard.cpp:182/203/230).  Pageable memory is copied to internal pinned
 * staging before the call returns.  Memory from ntc_host_alloc() is copied by
 * DMA directly and must stay untouched until ntc_wait(ticket) / ntc_sync().
 * Asynchronous and THREAD-SAFE: like ntRead, which the reference calls concurrently on one shared
 * sketch (one OpenMP thread per file, ntcard.cpp:445), every entry point that takes a context may be
 * called from several host threads; the context serialises them internally.
 * *ticket (optional) identifies the batch for ntc_wait(). */
int ntc_submit(ntc_ctx* ctx, const uint32_t* words, size_t n_words, const uint32_t* off, size_t n_rec,
    uint32_t stride_words, uint64_t* ticket);
/* Reads of ONE length (the common case: a sequencing run's 150 bp reads without N) need no length words: n_rec records of
 * ceil(len_bases / 16) words of bases each, back to back -- 40 instead of 44 bytes per 150 bp read over PCIe, which is what bounds
 * the end-to-end rate.  The length words are added on the device.  Otherwise like ntc_submit. */
int ntc_submit_bases(ntc_ctx* ctx, const uint32_t* bases, size_t n_rec, uint32_t len_bases, uint64_t* ticket);
/* Same, for a batch already resident in DEVICE memory (synthetic generator,
 * device-side producers).  The buffers must stay valid until ntc_sync().  Ragged batches (d_off != NULL) get the
 * checks of ntc_check_offsets on the device (NTC_EINVAL for broken offsets) and their longest record measured there,
 * which costs one synchronisation of the context's stream per call; uniform batches stay asynchronous. */
int ntc_submit_device(ntc_ctx* ctx, const uint32_t* d_words, size_t n_words, const uint32_t* d_off, size_t n_rec,
    uint32_t stride_words);
int ntc_wait(ntc_ctx* ctx, uint64_t ticket); /* host buffer of that batch is reusable */
int ntc_sync(ntc_ctx* ctx);                  /* all submitted work has finished (implies ntc_flush) */
/* The device path is a pipeline: batches are scanned and their sketch increments (ntComp's
 * ++t_Counter[...], ntcard.cpp:141-143) are first appended to a binned log in HBM; a flush applies the log
 * to the counters slice by slice while each slice is L2 resident.  The library flushes by itself when the
 * log fills up and inside every call that reads the counters (ntc_sync, ntc_finish, ntc_counters_device,
 * ntc_hist_range); ntc_flush forces one (asynchronous, ordered on the context's stream). */
int ntc_flush(ntc_ctx* ctx);

/* Multi-GPU plumbing: the device counters (uint32, [nK][2][2^rBits]; summed
 * across ranks by the caller's collective, then narrowed mod 2^16 by
 * ntc_finish) and the per-k totals (F1). */
int ntc_counters_device(ntc_ctx* ctx, void** d_counters, size_t* n_counters);
int ntc_totals(ntc_ctx* ctx, uint64_t* totKmer /* [nK] */); /* syncs */
int ntc_totals_nosync(ntc_ctx* ctx, uint64_t* totKmer);     /* the same, waiting for the stream but not flushing the hit log */
int ntc_set_totals(ntc_ctx* ctx, const uint64_t* totKmer);  /* after an all-reduce of F1 */

/* ---- multi-GPU: exchanging the hit log instead of the sketch ------------------------------------------
 * While nothing has been flushed, everything a rank has sampled sits in its hit log: counter indices in
 * blocks of 256, one list of blocks per SLICE of the sketch (slice s of n_slices covers counters
 * [s * counters_per_slice, (s+1) * counters_per_slice) of the flat [nK][2][2^rBits] array).  For inputs whose
 * hit count is small against the 2^(rBits+1) counters per k (10 M reads: 74 MB of log, 1 GiB of counters)
 * it is far cheaper to give every rank a subset of the slices, move the log blocks to their owners (one
 * all-to-all), and let each rank materialise and histogram only its own slices:
 *   ntc_log_counts   block count of every slice (syncs); *exportable = 0 when part of the log was already
 *                    flushed or added directly -- use the dense reduction (ntc_counters_device) then
 *   ntc_log_export   copy the blocks of `slices` (in that order) to d_blocks [sum nblk][260] uint32, a device
 *                    buffer of the caller: 256 entries, then the number of valid entries, then padding
 *   ntc_log_import   append blocks received from other ranks: runs[i] = {slice, n_blocks}, consecutive in
 *                    d_blocks; NTC_ENOMEM when they do not fit (nothing imported)
 *   ntc_flush_slices zero + apply only the slices with owned[s] != 0; the log of the others is dropped and
 *                    the context accepts only ntc_hist_slices / ntc_totals / ntc_reset afterwards
 *   ntc_hist_slices  counter-value histogram (v >= 1) of the owned slices, [nK][2][65536] uint32, to host
 *                    memory (p_hist, synchronous) and / or to a device buffer of the caller (d_p_hist,
 *                    asynchronous on the context's stream: ready for an all-reduce on the device)         */
int ntc_log_info(ntc_ctx* ctx, uint32_t* n_slices, uint64_t* counters_per_slice, uint32_t* entries_per_block);
/* pool_info (optional) [3]: blocks in use, blocks in the pool, block-list capacity per slice */
int ntc_log_counts(ntc_ctx* ctx, uint32_t* nblk /* [n_slices] */, int* exportable, uint32_t* pool_info);
int ntc_log_export(ntc_ctx* ctx, const uint32_t* slices, uint32_t n, void* d_blocks);
int ntc_log_import(ntc_ctx* ctx, const void* d_blocks, uint32_t n_blocks, const uint32_t* runs, uint32_t n_runs);
int ntc_flush_slices(ntc_ctx* ctx, const uint8_t* owned /* [n_slices] */);
/* ---- multi-GPU reduction over peer memory (NVLink) -----------------------------------------
 * One process per GPU on one node.  Replaces, across GPUs, what the reference gets for free from its single
 * shared sketch (ntcard.cpp:439, 142-143) and its F1 merge (ntcard.cpp:464-466).  Rank r owns the sketch slices
 * s with s % world == r.  Set-up, once: every rank exports the IPC handles of its hit log (ntc_peer_export),
 * the handles are all-gathered by the caller, every rank maps its peers' logs (ntc_peer_attach).  Per reduction,
 * all on the context's stream and without a host synchronisation:
 *   ntc_log_status_device(d_status)      d_status[0..nK) = F1 per k, d_status[nK] = 1 if this rank's log is no
 *                                        longer complete (flushed / direct increments), as int64
 *   <all-reduce (sum) of d_status>       F1 totals; also the barrier "every rank's log is complete"
 *   ntc_reduce_owned(d_status, d_hist)   zero the owned slices, apply to them the log entries of ALL ranks (remote
 *                                        logs are read through NVLink by the kernel itself), counter-value histogram
 *                                        of the owned slices -> d_hist [nK][2][65536] uint32 (values >= 1 only).
 *                                        Does nothing useful when d_status[nK] != 0: the caller then falls back to
 *                                        a dense reduction (ntc_counters_device + all-reduce).
 *   <all-reduce (sum) of d_hist>         the global histogram; also the barrier "nobody reads my log any more"
 * Afterwards only ntc_totals / ntc_set_totals / ntc_reset are valid (as after ntc_flush_slices).              */
#define NTC_PEER_HANDLE_BYTES 128
int ntc_peer_export(ntc_ctx* ctx, void* handles /* NTC_PEER_HANDLE_BYTES */);
int ntc_peer_attach(ntc_ctx* ctx, int world, int rank, const void* all_handles /* world x NTC_PEER_HANDLE_BYTES, by rank */);
/* the same for contexts of ONE process (several GPUs driven by one host program, or tests): all[rank] == ctx */
int ntc_peer_attach_contexts(ntc_ctx* ctx, int world, int rank, ntc_ctx* const* all);
int ntc_log_status_device(ntc_ctx* ctx, void* d_status /* int64 [nK + 1], device */);
int ntc_reduce_owned(ntc_ctx* ctx, const void* d_status, void* d_p_hist);
int ntc_stream_sync(ntc_ctx* ctx); /* wait for the context's stream WITHOUT flushing (before handing exported buffers to a collective) */
int ntc_hist_slices(ntc_ctx* ctx, const uint8_t* owned /* [n_slices] */, uint32_t* p_hist, void* d_p_hist);

/* Counter-value histogram (values >= 1 only; entry 0 of every table is left 0) of the range
 * [first, first+n) of the flat counter array, read from the DEVICE pointer d_counters (which points at
 * element `first`): the slice a rank owns after a reduce-scatter of the sketch.  Values are narrowed
 * mod 2^16 first.  p_hist: [nK][2][65536] uint32, host.  After summing the histograms of all slices the
 * caller sets p[t][0] = 2^rBits - sum_{v>=1} p[t][v].  first and n must be multiples of
 * min(65536, 2^rBits). */
int ntc_hist_range(ntc_ctx* ctx, const void* d_counters, uint64_t first, uint64_t n, uint32_t* p_hist);

/* Finish: wait for the device, narrow counters mod 2^16 (the reference's
 * uint16_t ++ wraps, ntcard.cpp:133,143) and return
 *   t_Counter : [nK][2][2^rBits] uint16, host, caller-owned, may be NULL
 *   totKmer   : [nK] = F1 (ntcard.cpp:155,466)
 *   p_hist    : [nK][2][65536] uint32 counter-value histogram computed on the
 *               device (ntcard.cpp:245-247), may be NULL. */
int ntc_finish(ntc_ctx* ctx, uint16_t* t_Counter, uint64_t* totKmer, uint32_t* p_hist);

/* Estimator, host side, plain double arithmetic in the reference's order
 * (compEst, ntcard.cpp:237-275), truncated at covMax (rows 1..covMax are
 * bit-identical to the untruncated recurrence).  Exactly one of p_hist
 * ([2][65536]) / t_Counter ([2][2^rBits]) is non-NULL.  f has covMax+1 entries;
 * f[0] = 0.  No device needed. */
int ntc_estimate(const uint32_t* p_hist, const uint16_t* t_Counter, unsigned rBits, unsigned sBits, unsigned covMax,
    double* F0, double* f);

/* ---- nthll: the HyperLogLog F0 estimator that ships beside ntcard -------------------------
 * Replaces, batch oriented like the sketch above:
 *   ntRead(const string& seq, uint8_t* mVec)      nthll.cpp:99-104   every canonical ntHash of the sequence ...
 *   ntComp(hVal, mVec)                            nthll.cpp:92-97    ... raises register hVal & (nBuck-1) to
 *                                                                    clz(hVal & ~(nBuck-1)) when those upper bits are not all 0
 *   tVec[j] = max over threads of mVec[j]         nthll.cpp:234-239
 *   the estimate main() prints                    nthll.cpp:243-254
 * ntc_hll_create makes a context in nthll mode (k = opt::kmLen, nBits = opt::nBits, nthll.cpp:45-48): batches go in
 * through the same ntc_submit / ntc_submit_device / ntc_wait / ntc_sync / ntc_reset / ntc_totals as above (the readers'
 * `seq.length() >= k` test, nthll.cpp:112,129,146, is implied: shorter records hold no k-mer); the entry points of the
 * ntCard sketch (ntc_finish, ntc_counters_device, ntc_log_*, ...) return NTC_ESTATE on it, and ntc_hll_* return
 * NTC_ESTATE on a sketch context.  d_regs: optional caller-owned DEVICE buffer of max(4, 2^nBits) bytes (e.g. a torch
 * uint8 tensor that torch.distributed will max-all-reduce), NULL lets the context allocate it.  The caller may change the
 * content of its buffer between two batches (never while one is in flight): the context looks at it afresh per batch. */
int ntc_hll_create(ntc_ctx** out, unsigned k, unsigned nBits, int device, void* d_regs, void* cuda_stream);
/* The registers on the device (uint8 [2^nBits]); multi-GPU: wait (ntc_sync), max-all-reduce them, then ntc_hll_finish. */
int ntc_hll_registers_device(ntc_ctx* ctx, void** d_regs, size_t* n_regs);
/* Wait for the device; regs: [2^nBits] uint8, host, may be NULL; totKmer: [1] number of k-mers hashed, may be NULL. */
int ntc_hll_finish(ntc_ctx* ctx, uint8_t* regs, uint64_t* totKmer);
/* Host side, no device needed: the reference's arithmetic in the reference's order (nthll.cpp:243-254); canon != 0
 * halves alpha (the reference always does, nthll.cpp:52,245).  nthll prints (unsigned long long)*est. */
int ntc_hll_estimate(const uint8_t* regs, unsigned nBits, int canon, double* est);

/* ---- host helpers ----------------------------------------------------------- */
/* Pinned host memory for zero-copy submission. */
void* ntc_host_alloc(size_t bytes);
void ntc_host_free(void* p);

/* Upper bound of words produced by ntc_pack_seqs for `n_seq` sequences holding
 * `total_bases` characters in total. */
size_t ntc_pack_bound(size_t n_seq, size_t total_bases);                       /* safe for any min_len (about one word per character) */
size_t ntc_pack_bound_k(size_t n_seq, size_t total_bases, uint32_t min_len); /* tighter: records shorter than min_len are dropped */
/* Split each sequence seq[i] = chars[seq_off[i] .. seq_off[i+1]) at characters
 * outside ACGTUacgtu, drop segments shorter than min_len, 2-bit pack the rest
 * as records.  Appends to words[*n_words..] / off[*n_rec..] (off gets
 * *n_rec+1 valid entries; off may be NULL).  Returns 0 or NTC_ENOMEM if
 * cap_words / cap_rec would be exceeded (nothing of the failing sequence is
 * kept; *consumed tells how many sequences were packed). */
int ntc_pack_seqs(const char* chars, const uint64_t* seq_off, size_t n_seq, uint32_t min_len, uint32_t* words,
    size_t cap_words, size_t* n_words, uint32_t* off, size_t cap_rec, size_t* n_rec, size_t* consumed);

/* The checks ntc_submit makes on the offsets of a ragged batch, for callers that build batches themselves: off[0..n_rec]
 * ascend strictly (every record has at least its length word) and off[n_rec] <= n_words.  *max_rec_words (optional) = words
 * of the longest record.  NTC_EINVAL names the first bad record.  No device needed. */
int ntc_check_offsets(const uint32_t* off, size_t n_rec, size_t n_words, uint32_t* max_rec_words);

/* Deterministic synthetic reads (mix64 generator).  mode 0:
 * uniform; mode 1: read i = gen(i mod U), reverse-complemented when (i/U) is
 * odd; mode 2: uniform with N runs (ASCII output only).
 * ntc_gen_ascii writes n*L characters; ntc_gen_packed writes n records of
 * stride_words words each (mode 0/1); ntc_gen_packed_device does the same on
 * the device, ordered on the context's stream (mode 0/1). */
int ntc_gen_ascii(uint64_t seed, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U, char* out);
int ntc_gen_packed(uint64_t seed, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U, uint32_t stride_words,
    uint32_t* words);
int ntc_gen_packed_device(ntc_ctx* ctx, uint64_t seed, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U,
    uint32_t stride_words, uint32_t* d_words);
uint32_t ntc_stride_words(uint32_t L, int align4); /* 1 + ceil(L/16), rounded up to 4 if align4 */

/* ---- introspection ----------------------------------------------------------- */
/* Counters for bench.py / tests: kernel launches issued by this context since
 * creation, and k-mers hashed as counted on the device. */
int ntc_stats(ntc_ctx* ctx, uint64_t* n_launches, uint64_t* n_batches);
/* Milliseconds the sketch kernels of the last ntc_sync()-ed batches took on the
 * device (CUDA events on the context's stream), and how many launches. */
int ntc_kernel_time(ntc_ctx* ctx, double* ms_total, uint64_t* n_timed);
/* The same split by pipeline stage, accumulated since the last call (after ntc_sync): ms3[0] scan kernels,
 * ms3[1] hit (+ fallback) kernels, ms3[2] apply kernels (flushes). */
int ntc_stage_times(ntc_ctx* ctx, double* ms3);
int ntc_device_count(void);
const char* ntc_last_error(void);
const char* ntc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NTCARD_B200_H */
