// reader.cpp -- see reader.h.
#include "reader.h"

#include <cctype>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <iostream>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace ntcb {

static bool ends_with(const std::string& s, const char* suf)
{
	size_t n = strlen(suf);
	return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

static std::string shell_quote(const std::string& s)
{
	std::string q = "'";
	for (char c : s) {
		if (c == '\'')
			q += "'\\''";
		else
			q += c;
	}
	return q + "'";
}

LineReader::LineReader(const std::string& path)
{
	const char* tool = nullptr;
	if (ends_with(path, ".gz") || ends_with(path, ".Z"))
		tool = "gunzip -c ";
	else if (ends_with(path, ".bz2"))
		tool = "bunzip2 -c ";
	else if (ends_with(path, ".xz"))
		tool = "xz -dc ";
	if (tool) {
		FILE* probe = fopen(path.c_str(), "rb");
		if (!probe)
			return;
		fclose(probe);
		fp_ = popen((std::string(tool) + shell_quote(path)).c_str(), "r");
		pipe_ = true;
	} else {
		fp_ = fopen(path.c_str(), "rb");
	}
	cap_ = (size_t)4 << 20;
	buf_ = (char*)malloc(cap_);
}

LineReader::~LineReader()
{
	if (fp_) {
		if (pipe_)
			pclose(fp_);
		else
			fclose(fp_);
	}
	free(buf_);
}

bool LineReader::fill()
{
	if (eof_)
		return false;
	if (beg_ > 0) {
		memmove(buf_, buf_ + beg_, end_ - beg_);
		end_ -= beg_;
		beg_ = 0;
	}
	if (end_ == cap_) { // a line longer than the buffer (FASTA chromosomes on one line)
		cap_ *= 2;
		buf_ = (char*)realloc(buf_, cap_);
	}
	size_t n = fread(buf_ + end_, 1, cap_ - end_, fp_);
	if (n == 0) {
		eof_ = true;
		return false;
	}
	end_ += n;
	return true;
}

bool LineReader::next(const char** line, size_t* len)
{
	if (!fp_)
		return false;
	size_t scanned = 0;
	for (;;) {
		char* nl = (char*)memchr(buf_ + beg_ + scanned, '\n', end_ - beg_ - scanned);
		if (nl) {
			*line = buf_ + beg_;
			*len = (size_t)(nl - (buf_ + beg_));
			beg_ = (size_t)(nl - buf_) + 1;
			return true;
		}
		scanned = end_ - beg_;
		if (!fill()) {
			if (end_ > beg_) { // last line without a trailing newline
				*line = buf_ + beg_;
				*len = end_ - beg_;
				beg_ = end_;
				return true;
			}
			return false;
		}
	}
}

// ------------------------------------------------------------------------------------------------
BatchSubmitter::BatchSubmitter(ntc_ctx* ctx, unsigned min_len, std::mutex* submit_mu, size_t words_per_buffer)
    : ctx_(ctx), min_len_(min_len), mu_(submit_mu)
{
	init(words_per_buffer, true);
}

BatchSubmitter::BatchSubmitter(std::function<ntc_ctx*()> ctx_source, unsigned min_len, std::mutex* submit_mu, size_t words_per_buffer)
    : ctx_(nullptr), ctx_source_(std::move(ctx_source)), min_len_(min_len), mu_(submit_mu)
{
	init(words_per_buffer, false); // pageable: pinning would wait for the CUDA initialisation this constructor is meant to overlap
}

void BatchSubmitter::init(size_t words_per_buffer, bool pinned)
{
	for (auto& b : uni_.buf)
		alloc(b, words_per_buffer, 1, pinned);
	for (auto& b : rag_.buf)
		alloc(b, words_per_buffer / 4, words_per_buffer / 16, pinned);
	tmp_words_.resize(1 << 16);
	tmp_off_.resize((1 << 14) + 1);
}

ntc_ctx* BatchSubmitter::ctx()
{
	if (!ctx_ && ctx_source_)
		ctx_ = ctx_source_(); // blocks until the context exists
	return ctx_;
}

BatchSubmitter::~BatchSubmitter()
{
	for (auto& b : uni_.buf)
		release(b);
	for (auto& b : rag_.buf)
		release(b);
}

static void die_ntc()
{
	std::cerr << "ntCard: " << ntc_last_error() << "\n";
	exit(EXIT_FAILURE);
}

void BatchSubmitter::alloc(Buf& b, size_t words, size_t recs, bool pinned)
{
	if (pinned) {
		b.words = (uint32_t*)ntc_host_alloc(words * sizeof(uint32_t));
		b.off = (uint32_t*)ntc_host_alloc((recs + 1) * sizeof(uint32_t));
	} else {
		b.words = (uint32_t*)malloc(words * sizeof(uint32_t));
		b.off = (uint32_t*)malloc((recs + 1) * sizeof(uint32_t));
	}
	if (!b.words || !b.off) {
		std::cerr << "ntCard: cannot allocate host memory: " << (pinned ? ntc_last_error() : "malloc failed") << "\n";
		exit(EXIT_FAILURE);
	}
	b.pinned = pinned;
	b.submits = 0;
	b.cap_words = words;
	b.cap_rec = recs;
	b.n_words = b.n_rec = 0;
	b.ticket = 0;
}

void BatchSubmitter::release(Buf& b)
{
	if (b.pinned) {
		ntc_host_free(b.words);
		ntc_host_free(b.off);
	} else {
		free(b.words);
		free(b.off);
	}
	b.words = b.off = nullptr;
}

void BatchSubmitter::flush_stream(Stream& st)
{
	struct Timer {
		double* acc;
		std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
		~Timer() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
	} timer{ &submit_seconds_ };
	Buf& b = st.buf[st.cur];
	if (b.n_rec > 0) {
		ntc_ctx* c = ctx();
		{
			std::lock_guard<std::mutex> lk(*mu_);
			const int rc = st.stride ? ntc_submit(c, b.words, b.n_words, NULL, b.n_rec, st.stride, &b.ticket)
			                         : ntc_submit(c, b.words, b.n_words, b.off, b.n_rec, 0, &b.ticket);
			if (rc)
				die_ntc();
		}
		if (!b.pinned) {
			b.ticket = 0; // pageable memory was copied to the library's staging before ntc_submit returned
			if (++b.submits >= 4) { // a buffer that keeps being reused is worth pinning (DMA straight from it, no staging copy)
				const size_t w = b.cap_words, r = b.cap_rec;
				release(b);
				alloc(b, w, r, true);
			}
		}
	}
	st.cur ^= 1;
	Buf& n = st.buf[st.cur];
	if (n.ticket) { // the DMA engine may still be reading the buffer we are about to refill
		std::lock_guard<std::mutex> lk(*mu_);
		if (ntc_wait(ctx_, n.ticket))
			die_ntc();
		n.ticket = 0;
	}
	n.n_words = n.n_rec = 0;
}

void BatchSubmitter::append(Stream& st, const uint32_t* rec, size_t nwords)
{
	const size_t need = st.stride ? st.stride : nwords;
	for (int attempt = 0; attempt < 2; attempt++) {
		Buf& b = st.buf[st.cur];
		if (b.n_words + need <= b.cap_words && (st.stride || b.n_rec + 1 <= b.cap_rec)) {
			memcpy(b.words + b.n_words, rec, nwords * sizeof(uint32_t));
			if (need > nwords)
				memset(b.words + b.n_words + nwords, 0, (need - nwords) * sizeof(uint32_t));
			if (!st.stride) {
				b.off[b.n_rec] = (uint32_t)b.n_words;
				b.off[b.n_rec + 1] = (uint32_t)(b.n_words + need);
			}
			b.n_words += need;
			b.n_rec++;
			return;
		}
		if (b.n_rec > 0) {
			flush_stream(st);
			continue;
		}
		// a single record larger than an empty buffer (e.g. a chromosome): grow this buffer
		if (b.ticket && ntc_wait(ctx_, b.ticket))
			die_ntc();
		const bool was_pinned = b.pinned;
		release(b);
		alloc(b, need + (need >> 2), b.cap_rec, was_pinned);
	}
	std::cerr << "ntCard: internal error: record does not fit the batch buffer\n";
	exit(EXIT_FAILURE);
}

void BatchSubmitter::add(const char* seq, size_t len)
{
	if (len < min_len_)
		return; // no window fits (ntHashIterator.hpp:61-64)
	const size_t bound = ntc_pack_bound_k(1, len, min_len_);
	if (tmp_words_.size() < bound)
		tmp_words_.resize(bound);
	const size_t rec_bound = len / (min_len_ ? min_len_ : 1) + 2;
	if (tmp_off_.size() < rec_bound + 1)
		tmp_off_.resize(rec_bound + 1);
	const uint64_t so[2] = { 0, (uint64_t)len };
	size_t nw = 0, nr = 0, consumed = 0;
	if (ntc_pack_seqs(seq, so, 1, min_len_, tmp_words_.data(), tmp_words_.size(), &nw, tmp_off_.data(), tmp_off_.size() - 1, &nr, &consumed))
		die_ntc();
	for (size_t i = 0; i < nr; i++) {
		const uint32_t* rec = tmp_words_.data() + tmp_off_[i];
		const size_t nwords = tmp_off_[i + 1] - tmp_off_[i];
		const uint32_t padded = (uint32_t)((nwords + 3) & ~(size_t)3);
		if (uni_.stride == 0 && padded <= 64)
			uni_.stride = padded; // the first record fixes the common size (reads of one run share a length)
		if (padded == uni_.stride)
			append(uni_, rec, nwords);
		else
			append(rag_, rec, nwords);
	}
}

void BatchSubmitter::flush()
{
	flush_stream(uni_);
	flush_stream(rag_);
}

void BatchSubmitter::finish()
{
	std::lock_guard<std::mutex> lk(*mu_);
	for (Stream* st : { &uni_, &rag_ })
		for (auto& b : st->buf)
			if (b.ticket) {
				ntc_wait(ctx_, b.ticket);
				b.ticket = 0;
			}
}

// ------------------------------------------------------------------------------------------------
static bool all_digits(const std::string& s) // isNumber, ntcard.cpp:96-103
{
	if (s.empty())
		return false;
	for (unsigned char c : s)
		if (!isdigit(c))
			return false;
	return true;
}

// whitespace-separated token `idx` (0-based) of a line; false if the line has fewer tokens
static bool token_at(const char* line, size_t len, unsigned idx, const char** tok, size_t* tlen)
{
	size_t i = 0;
	for (unsigned t = 0;; t++) {
		while (i < len && isspace((unsigned char)line[i]))
			i++;
		if (i >= len)
			return false;
		size_t j = i;
		while (j < len && !isspace((unsigned char)line[j]))
			j++;
		if (t == idx) {
			*tok = line + i;
			*tlen = j - i;
			return true;
		}
		i = j;
	}
}

// FASTQ body after the first header was consumed by the sniffer (getEfq, ntcard.cpp:173-189): a
// record is hashed only when its quality line could be read.
static void read_fastq(LineReader& in, BatchSubmitter& sub)
{
	std::string seq;
	const char* l;
	size_t n;
	for (;;) {
		if (!in.next(&l, &n))
			return;
		seq.assign(l, n);
		if (!in.next(&l, &n)) // '+'
			return;
		if (!in.next(&l, &n)) // quality
			return;
		sub.add(seq.data(), seq.size());
		if (!in.next(&l, &n)) // next header
			return;
	}
}

// FASTA (getEfa, ntcard.cpp:191-208): lines up to the next '>' are concatenated into one sequence.
static void read_fasta(LineReader& in, BatchSubmitter& sub)
{
	std::string seq;
	const char* l;
	size_t n;
	bool good = true;
	while (good) {
		seq.clear();
		good = in.next(&l, &n);
		while (good && !(n > 0 && l[0] == '>')) {
			seq.append(l, n);
			good = in.next(&l, &n);
		}
		sub.add(seq.data(), seq.size());
	}
}

// SAM (getEsm, ntcard.cpp:210-235): column 10 of every alignment line.  `seq` persists across
// lines: a line with fewer than 10 fields leaves it unchanged and the previous sequence is hashed
// again, exactly as the reference's `iss >> ... >> seq` does.
static void read_sam(LineReader& in, BatchSubmitter& sub, bool has_header, const std::string& first_line)
{
	std::string line, seq;
	const char* l;
	size_t n;
	if (has_header) {
		bool got = false;
		while (in.next(&l, &n)) {
			if (!(n > 0 && l[0] == '@')) {
				line.assign(l, n);
				got = true;
				break;
			}
		}
		if (!got)
			return;
	} else {
		line = first_line;
	}
	for (;;) {
		const char* tok;
		size_t tlen;
		if (token_at(line.data(), line.size(), 9, &tok, &tlen))
			seq.assign(tok, tlen);
		sub.add(seq.data(), seq.size());
		if (!in.next(&l, &n))
			return;
		line.assign(l, n);
	}
}

bool read_file(const std::string& path, BatchSubmitter& sub, bool nthll_rules)
{
	LineReader in(path);
	if (!in.ok())
		return nthll_rules; // nthll never checks the stream: an unreadable file contributes nothing (nthll.cpp:219-229)
	const char* l;
	size_t n;
	std::string h;
	if (in.next(&l, &n))
		h.assign(l, n);
	// getftype, ntcard.cpp:105-130
	if (!h.empty() && h[0] == '>') {
		read_fasta(in, sub);
		return true;
	}
	if (!h.empty() && h[0] == '@') {
		const char a = h.size() > 1 ? h[1] : 0, b = h.size() > 2 ? h[2] : 0;
		if ((a == 'H' && b == 'D') || (a == 'S' && b == 'Q') || (a == 'R' && b == 'G') || (a == 'P' && b == 'G') ||
		    (a == 'C' && b == 'O')) {
			read_sam(in, sub, true, h);
		} else {
			read_fastq(in, sub);
		}
		return true;
	}
	const char *t2, *t5;
	size_t n2, n5;
	if (token_at(h.data(), h.size(), 1, &t2, &n2) && token_at(h.data(), h.size(), 4, &t5, &n5) &&
	    all_digits(std::string(t2, n2)) && all_digits(std::string(t5, n5))) {
		read_sam(in, sub, false, h);
		return true;
	}
	if (nthll_rules) { // nthll's getftype has no "unknown" answer: everything else is header-less SAM (nthll.cpp:86-88)
		read_sam(in, sub, false, h);
		return true;
	}
	return false;
}

// ------------------------------------------------------------------------------------------------
namespace {
struct Piece {
	size_t a, b;          // byte range
	size_t newlines = 0;  // '\n' in [a, b)
	size_t first_line = 0; // number of newlines before a = index of the line that contains byte a
};

// FASTQ: sequence lines are the lines with index % 4 == 1 (line 0 = the header the sniffer consumed); hashed iff line index + 2 exists
void parse_fastq_piece(const char* d, size_t size, const Piece& pc, size_t n_lines, BatchSubmitter& sub)
{
	size_t p = pc.a, idx = pc.first_line;
	if (p > 0 && d[p - 1] != '\n') { // the line containing a started in the previous piece: skip to the next line start
		const char* nl = (const char*)memchr(d + p, '\n', size - p);
		if (!nl)
			return;
		p = (size_t)(nl - d) + 1;
		idx++;
	}
	while (p < pc.b && p < size) {
		const char* nl = (const char*)memchr(d + p, '\n', size - p);
		const size_t e = nl ? (size_t)(nl - d) : size;
		if ((idx & 3u) == 1u && idx + 2 < n_lines)
			sub.add(d + p, e - p);
		if (!nl)
			return;
		p = e + 1;
		idx++;
	}
}

// FASTA: a record belongs to the piece in which its '>' line starts; its sequence is every following line up to the next '>' line
void parse_fasta_piece(const char* d, size_t size, const Piece& pc, BatchSubmitter& sub)
{
	size_t p = pc.a;
	if (p > 0 && d[p - 1] != '\n') {
		const char* nl = (const char*)memchr(d + p, '\n', size - p);
		if (!nl)
			return;
		p = (size_t)(nl - d) + 1;
	}
	std::string seq;
	bool in_record = false; // inside a record whose header line started in this piece
	while (p < size) {
		const char* nl = (const char*)memchr(d + p, '\n', size - p);
		const size_t e = nl ? (size_t)(nl - d) : size;
		if (e > p && d[p] == '>') {
			if (in_record)
				sub.add(seq.data(), seq.size());
			if (p >= pc.b) // the next piece's record
				return;
			seq.clear();
			in_record = true;
		} else if (in_record) {
			seq.append(d + p, e - p);
		}
		if (!nl)
			break;
		p = e + 1;
	}
	if (in_record)
		sub.add(seq.data(), seq.size());
}
} // namespace

bool read_file_parallel(const std::string& path, const std::function<ntc_ctx*()>& ctx_source, unsigned min_len, std::mutex* submit_mu,
    unsigned nthreads, bool* handled)
{
	*handled = false;
	if (nthreads < 2 || ends_with(path, ".gz") || ends_with(path, ".Z") || ends_with(path, ".bz2") || ends_with(path, ".xz"))
		return true;
	const int fd = open(path.c_str(), O_RDONLY);
	if (fd < 0)
		return true; // read_file reports the error the reference's way
	struct stat st;
	if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || (size_t)st.st_size < ((size_t)8 << 20) * nthreads) {
		close(fd);
		return true;
	}
	const size_t size = (size_t)st.st_size;
	const char* d = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (d == MAP_FAILED)
		return true;
	madvise((void*)d, size, MADV_SEQUENTIAL);
	// getftype (ntcard.cpp:105-130) on the first line: only FASTA and FASTQ are cut up; SAM and anything else go the serial way
	int type = -1;
	if (d[0] == '>') {
		type = 1;
	} else if (d[0] == '@') {
		const char a = size > 1 ? d[1] : 0, b = size > 2 ? d[2] : 0;
		const bool sam = (a == 'H' && b == 'D') || (a == 'S' && b == 'Q') || (a == 'R' && b == 'G') || (a == 'P' && b == 'G') || (a == 'C' && b == 'O');
		if (!sam)
			type = 0;
	}
	if (type < 0) {
		munmap((void*)d, size);
		return true;
	}
	*handled = true;
	std::vector<Piece> pc(nthreads);
	for (unsigned t = 0; t < nthreads; t++) {
		pc[t].a = size / nthreads * t;
		pc[t].b = t + 1 == nthreads ? size : size / nthreads * (t + 1);
	}
	std::vector<std::thread> th;
	for (unsigned t = 0; t < nthreads; t++) // pass 1: newlines per piece
		th.emplace_back([&, t]() {
			size_t n = 0;
			const char* p = d + pc[t].a;
			const char* const e = d + pc[t].b;
			while (p < e) {
				const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
				if (!nl)
					break;
				n++;
				p = nl + 1;
			}
			pc[t].newlines = n;
		});
	for (auto& x : th)
		x.join();
	th.clear();
	size_t total_nl = 0;
	for (unsigned t = 0; t < nthreads; t++) {
		pc[t].first_line = total_nl;
		total_nl += pc[t].newlines;
	}
	const size_t n_lines = total_nl + (d[size - 1] != '\n' ? 1 : 0); // a last line without '\n' counts (std::getline)
	for (unsigned t = 0; t < nthreads; t++) // pass 2: parse + pack + submit
		th.emplace_back([&, t]() {
			BatchSubmitter sub(ctx_source, min_len, submit_mu);
			if (type == 0)
				parse_fastq_piece(d, size, pc[t], n_lines, sub);
			else
				parse_fasta_piece(d, size, pc[t], sub);
			sub.flush();
			sub.finish();
		});
	for (auto& x : th)
		x.join();
	munmap((void*)d, size);
	return true;
}

} // namespace ntcb
