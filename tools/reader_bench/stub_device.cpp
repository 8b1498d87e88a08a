// stub_device.cpp -- TEST/BENCH INFRASTRUCTURE ONLY (never linked into the product): the device entry points the CLI's reader
// threads call, as no-ops, so that the host ingest (line reader -> format parsers -> N-split packer -> batch buffers) can be
// timed on a machine without a GPU.  ntc_pack_seqs / ntc_pack_bound are the real ones (host_util.cpp is linked in).
#include <atomic>
#include <cstdlib>

#include "ntcard_b200.h"

namespace ntc {
const char* last_err(); // host_util.cpp
}

static std::atomic<uint64_t> g_ticket{ 1 }, g_words{ 0 }, g_recs{ 0 }, g_batches{ 0 }, g_digest{ 0 };

// order-independent digest of the submitted records: sum over records of an FNV-1a hash of (length word, base words)
static uint64_t record_hash(const uint32_t* r)
{
	const uint32_t nw = 1 + (r[0] + 15) / 16;
	uint64_t h = 0xcbf29ce484222325ull;
	for (uint32_t i = 0; i < nw; i++)
		h = (h ^ r[i]) * 0x100000001b3ull;
	return h;
}

extern "C" {
int ntc_submit(ntc_ctx*, const uint32_t* words, size_t n_words, const uint32_t* off, size_t n_rec, uint32_t stride, uint64_t* ticket)
{
	uint64_t d = 0;
	for (size_t i = 0; i < n_rec; i++)
		d += record_hash(words + (off ? off[i] : i * stride));
	g_digest += d;
	g_words += n_words;
	g_recs += n_rec;
	g_batches++;
	if (ticket)
		*ticket = g_ticket++;
	return 0;
}
int ntc_wait(ntc_ctx*, uint64_t) { return 0; }
const char* ntc_last_error(void) { return ntc::last_err(); }
void* ntc_host_alloc(size_t bytes) { return malloc(bytes); }
void ntc_host_free(void* p) { free(p); }
uint64_t stub_words() { return g_words; }
uint64_t stub_recs() { return g_recs; }
uint64_t stub_batches() { return g_batches; }
uint64_t stub_digest() { return g_digest; }
}
