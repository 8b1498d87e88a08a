// reader.h -- host ingest of the ntcard CLI: FASTQ / FASTA / SAM parsing with the reference's
// record rules (ntcard.cpp:105-130 getftype, 173-189 getEfq, 191-208 getEfa, 210-235 getEsm), then
// N-split + 2-bit packing into pinned double buffers and batch submission through the C-ABI.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "ntcard_b200.h"

namespace ntcb {

// Buffered line reader over a file or a decompressor pipe (gz/bz2/xz by extension, like the
// reference's fopen hook, Common/Uncompress.cpp:46-67).  next() has std::getline semantics: returns
// false only at EOF with nothing read; the line excludes the '\n' (a '\r' is kept).
class LineReader {
public:
	explicit LineReader(const std::string& path);
	~LineReader();
	bool ok() const { return fp_ != nullptr; }
	bool next(const char** line, size_t* len);

private:
	bool fill();
	FILE* fp_ = nullptr;
	bool pipe_ = false;
	char* buf_ = nullptr;
	size_t cap_ = 0, beg_ = 0, end_ = 0;
	bool eof_ = false;
};

// Packs sequences into pinned buffers and submits full ones (asynchronously).  Records are routed into two
// streams, each double buffered: records whose padded size equals that of the first record seen (the common
// read length, e.g. 150 bp -> 12 words) go into a UNIFORM-STRIDE batch, which the device runs through the
// bit-sliced kernel; everything else (segments cut short by an N, odd lengths) goes into a ragged batch.
// The context may still be under construction when the readers start (creating it -- CUDA initialisation -- takes longer than
// parsing a few hundred MB): with a context SOURCE the submitter starts on pageable buffers, asks for the context only when its
// first batch is full (blocking until it exists), and moves a buffer to pinned memory once it has been submitted a few times.
class BatchSubmitter {
public:
	BatchSubmitter(ntc_ctx* ctx, unsigned min_len, std::mutex* submit_mu, size_t words_per_buffer = (size_t)2 << 20); // 8 MB per pinned buffer: pinning is slow (~1 GB/s), 16 reader threads x 84 MB cost 1 s
	BatchSubmitter(std::function<ntc_ctx*()> ctx_source, unsigned min_len, std::mutex* submit_mu, size_t words_per_buffer = (size_t)2 << 20);
	~BatchSubmitter();
	void add(const char* seq, size_t len); // == one ntRead(seq, ...) call of the reference
	void flush();                           // submit what is buffered (both streams)
	double seconds_in_submit() const { return submit_seconds_; } // time spent in ntc_submit / ntc_wait incl. the lock
	void finish();                          // wait until the device no longer needs our buffers

private:
	double submit_seconds_ = 0;
	struct Buf {
		uint32_t* words = nullptr;
		uint32_t* off = nullptr;
		size_t cap_words = 0, cap_rec = 0;
		size_t n_words = 0, n_rec = 0;
		uint64_t ticket = 0;
		bool pinned = false;
		unsigned submits = 0;
	};
	struct Stream {
		Buf buf[2];
		int cur = 0;
		uint32_t stride = 0; // > 0: uniform-stride stream
	};
	void alloc(Buf& b, size_t words, size_t recs, bool pinned);
	void release(Buf& b);
	void init(size_t words_per_buffer, bool pinned);
	ntc_ctx* ctx();
	void flush_stream(Stream& st);
	void append(Stream& st, const uint32_t* rec, size_t nwords);
	ntc_ctx* ctx_;
	std::function<ntc_ctx*()> ctx_source_;
	unsigned min_len_;
	std::mutex* mu_;
	Stream uni_, rag_;
	std::vector<uint32_t> tmp_words_, tmp_off_;
};

// Sniff the format on the first line and feed every sequence of the file to `sub`.
// Returns false when the format is not recognised / the file cannot be read (ntcard.cpp:459-462).
// nthll_rules: the sniffer of nthll.cpp:73-90 instead -- the same three formats, but whatever is not FASTA / FASTQ /
// SAM-with-header is read as header-less SAM, and unreadable files are skipped silently; never returns false.
bool read_file(const std::string& path, BatchSubmitter& sub, bool nthll_rules = false);

// One large uncompressed FASTQ / FASTA file parsed by `nthreads` threads (the reference reads a file with ONE thread however
// many -t says, ntcard.cpp:445).  The file is mapped and cut at arbitrary byte offsets; a first pass counts the newlines of
// every piece in parallel, which gives every piece the index of its first line -- so the FASTQ rule "lines 1, 5, 9, ... are
// sequences, hashed only if their quality line exists" (getEfq, ntcard.cpp:173-189) is applied exactly, with no guessing at
// '@' / '+' characters; FASTA records (getEfa, ntcard.cpp:191-208) belong to the piece their '>' line starts in.  Every
// thread has its own BatchSubmitter; the sketch does not depend on the order of the sequences.
// *handled = false (and nothing read) when the file is compressed, small, SAM or unrecognised: use read_file then.
bool read_file_parallel(const std::string& path, const std::function<ntc_ctx*()>& ctx_source, unsigned min_len, std::mutex* submit_mu,
    unsigned nthreads, bool* handled);

} // namespace ntcb
