// host_util.cpp -- host-side parts of the C-ABI that need no device: error string, the N-split
// 2-bit packer, the synthetic read generator, and the estimator (compEst).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/types.h>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/ntcard_b200.h"
#include "internal.h"
#include "nthash_device.cuh"

namespace ntc {

static thread_local char g_err[512] = "";

int set_err(int code, const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}

const char* last_err() { return g_err; }

// 2-bit code of a character, 4 = invalid.  Valid set = the characters with a nonzero seedTab entry
// among printable text: A C G T U in either case (nthash.hpp:31-64).  Deliberate deviation: the
// control bytes 0x01,0x03,0x04,0x05,0x07 also carry seeds in the reference's table (row 0..7 serves
// the `c & cpOff` complement lookup, nthash.hpp:16,32); they never occur in sequence text and are
// treated as invalid here.
static const struct CodeTab {
	uint8_t t[256];
	CodeTab()
	{
		memset(t, 4, sizeof t);
		t[(unsigned char)'A'] = t[(unsigned char)'a'] = 0;
		t[(unsigned char)'C'] = t[(unsigned char)'c'] = 1;
		t[(unsigned char)'G'] = t[(unsigned char)'g'] = 2;
		t[(unsigned char)'T'] = t[(unsigned char)'t'] = 3;
		t[(unsigned char)'U'] = t[(unsigned char)'u'] = 3;
	}
} g_code;

// ---- 16 characters at a time (SSSE3; the packer runs on the submitting / reader threads and bounds the command line) ----
// valid16: bit q set = character q is one of ACGTUacgtu.  pack16: the 2-bit codes of 16 valid characters, base q in bits 2q..2q+1.
// The code comes straight from the ASCII bits: (c >> 1) & 3 is A=0 C=1 T/U=2 G=3 in either case; x ^ (x >> 1) swaps 2 and 3.
#if defined(__x86_64__)
#define NTC_HAVE_SIMD_PACK 1
__attribute__((target("ssse3"))) static inline unsigned valid16(const unsigned char* p)
{
	const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
	const __m128i up = _mm_and_si128(v, _mm_set1_epi8((char)0xDF)); // fold case
	__m128i ok = _mm_cmpeq_epi8(up, _mm_set1_epi8('A'));
	ok = _mm_or_si128(ok, _mm_cmpeq_epi8(up, _mm_set1_epi8('C')));
	ok = _mm_or_si128(ok, _mm_cmpeq_epi8(up, _mm_set1_epi8('G')));
	ok = _mm_or_si128(ok, _mm_cmpeq_epi8(up, _mm_set1_epi8('T')));
	ok = _mm_or_si128(ok, _mm_cmpeq_epi8(up, _mm_set1_epi8('U')));
	return (unsigned)_mm_movemask_epi8(ok);
}

__attribute__((target("ssse3"))) static inline uint32_t pack16(const unsigned char* p)
{
	const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
	__m128i c = _mm_and_si128(_mm_srli_epi16(v, 1), _mm_set1_epi8(3));
	c = _mm_xor_si128(c, _mm_and_si128(_mm_srli_epi16(c, 1), _mm_set1_epi8(1)));
	const __m128i t = _mm_maddubs_epi16(c, _mm_set1_epi16(0x0401));  // byte pairs  -> 4-bit values in 16-bit lanes
	const __m128i u = _mm_madd_epi16(t, _mm_set1_epi32(0x00100001)); // lane pairs  -> 8-bit values in 32-bit lanes
	const __m128i b = _mm_shuffle_epi8(u, _mm_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
	return (uint32_t)_mm_cvtsi128_si32(b);
}

static const bool g_simd_pack = __builtin_cpu_supports("ssse3");
#else
#define NTC_HAVE_SIMD_PACK 0
static const bool g_simd_pack = false;
#endif

static inline uint64_t mix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ULL;
	uint64_t z = x;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

// 2-bit codes of read i (length L) into out[0..L)
static void gen_codes(uint64_t S, uint64_t i, uint32_t L, int mode, uint64_t U, uint8_t* out)
{
	const bool rep = (mode == 1 && U != 0);
	const uint64_t src = rep ? i % U : i;
	for (uint32_t w = 0; w * 32 < L; w++) {
		uint64_t v = mix64((S << 48) ^ (src << 12) ^ (uint64_t)w);
		for (uint32_t j = 0; j < 32 && w * 32 + j < L; j++)
			out[w * 32 + j] = (uint8_t)((v >> (2 * j)) & 3);
	}
	if (rep && ((i / U) & 1)) {
		for (uint32_t a = 0, b = L - 1; a < b; a++, b--) {
			uint8_t t = out[a];
			out[a] = out[b];
			out[b] = t;
		}
		for (uint32_t a = 0; a < L; a++)
			out[a] = 3 - out[a];
	}
}

} // namespace ntc

using ntc::set_err;

extern "C" {

uint32_t ntc_stride_words(uint32_t L, int align4)
{
	uint32_t w = 1 + (L + 15) / 16;
	if (align4)
		w = (w + 3) & ~3u;
	return w;
}

size_t ntc_pack_bound_k(size_t n_seq, size_t total_bases, uint32_t min_len)
{
	// A record of len >= min_len bases costs 1 length word + ceil(len/16) <= 2 + len/16 words.  Every record but the last of a
	// sequence is followed by at least one invalid character, so a sequence of c characters yields at most c/(min_len+1) + 1
	// records: words <= total/16 + 2 * (n_seq + total/(min_len+1)).  min_len = 1 ("ANAN..."): about one word per character.
	const size_t m = min_len ? min_len : 1;
	return total_bases / 16 + 2 * (n_seq + total_bases / (m + 1)) + 2;
}

size_t ntc_pack_bound(size_t n_seq, size_t total_bases) { return ntc_pack_bound_k(n_seq, total_bases, 1); }

int ntc_pack_seqs(const char* chars, const uint64_t* seq_off, size_t n_seq, uint32_t min_len, uint32_t* words, size_t cap_words,
    size_t* n_words, uint32_t* off, size_t cap_rec, size_t* n_rec, size_t* consumed)
{
	if (!chars || !seq_off || !words || !n_words || !n_rec)
		return set_err(NTC_EINVAL, "ntc_pack_seqs: null argument");
	size_t nw = *n_words, nr = *n_rec;
	if (consumed)
		*consumed = 0;
	if (min_len < 1)
		min_len = 1;
	for (size_t s = 0; s < n_seq; s++) {
		const unsigned char* p = (const unsigned char*)chars + seq_off[s];
		const uint64_t L = seq_off[s + 1] - seq_off[s];
		const size_t nw0 = nw, nr0 = nr;
		uint64_t i = 0;
		bool overflow = false;
		while (i < L) {
			while (i < L && ntc::g_code.t[p[i]] > 3)
				i++;
			uint64_t j = i;
#if NTC_HAVE_SIMD_PACK
			if (ntc::g_simd_pack)
				while (j + 16 <= L) { // whole blocks of valid characters
					const unsigned m = ntc::valid16(p + j);
					if (m != 0xFFFFu) {
						j += (unsigned)__builtin_ctz(~m); // first invalid character of the block ends the run
						goto run_end;
					}
					j += 16;
				}
#endif
			while (j < L && ntc::g_code.t[p[j]] <= 3)
				j++;
#if NTC_HAVE_SIMD_PACK
		run_end:;
#endif
			const uint64_t len = j - i;
			if (len >= min_len) {
				if (len > 0xFFFFFFFFull)
					return set_err(NTC_EINVAL, "ntc_pack_seqs: segment longer than 2^32-1 bases");
				const size_t need = 1 + (size_t)((len + 15) / 16);
				if (nw + need > cap_words || nr + 1 > cap_rec || nw + need > 0xFFFFFFF0ull) {
					overflow = true;
					break;
				}
				if (off)
					off[nr] = (uint32_t)nw;
				words[nw++] = (uint32_t)len;
				uint64_t b = i;
#if NTC_HAVE_SIMD_PACK
				if (ntc::g_simd_pack)
					for (; b + 16 <= j; b += 16)
						words[nw++] = ntc::pack16(p + b);
#endif
				for (; b + 16 <= j; b += 16) {
					uint32_t w = 0;
					for (unsigned q = 0; q < 16; q++)
						w |= (uint32_t)ntc::g_code.t[p[b + q]] << (2 * q);
					words[nw++] = w;
				}
				if (b < j) {
					uint32_t w = 0;
#if NTC_HAVE_SIMD_PACK
					if (ntc::g_simd_pack && b + 16 <= L) // the 16 characters exist: pack them all, keep the run's
						w = ntc::pack16(p + b) & ((1u << (2 * (unsigned)(j - b))) - 1u);
					else
#endif
						for (unsigned q = 0; b + q < j; q++)
							w |= (uint32_t)ntc::g_code.t[p[b + q]] << (2 * q);
					words[nw++] = w;
				}
				nr++;
			}
			i = j;
		}
		if (overflow) {
			nw = nw0;
			nr = nr0;
			if (off)
				off[nr] = (uint32_t)nw;
			*n_words = nw;
			*n_rec = nr;
			return set_err(NTC_ENOMEM, "ntc_pack_seqs: output buffers full after %zu sequences", s);
		}
		if (consumed)
			*consumed = s + 1;
	}
	if (off)
		off[nr] = (uint32_t)nw;
	*n_words = nw;
	*n_rec = nr;
	return NTC_OK;
}

int ntc_gen_ascii(uint64_t seed, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U, char* out)
{
	if (!out || mode < 0 || mode > 2)
		return set_err(NTC_EINVAL, "ntc_gen_ascii: bad argument");
	static const char B[4] = { 'A', 'C', 'G', 'T' };
	std::vector<uint8_t> codes(L);
	for (uint64_t r = 0; r < n; r++) {
		const uint64_t i = first + r;
		ntc::gen_codes(seed, i, L, mode, U, codes.data());
		char* o = out + r * (uint64_t)L;
		for (uint32_t j = 0; j < L; j++)
			o[j] = B[codes[j]];
		if (mode == 2) {
			uint64_t nn = ntc::mix64((seed << 48) ^ (i << 12) ^ 0xFFF);
			unsigned runs = (unsigned)(nn & 3);
			for (unsigned t = 0; t < runs; t++) {
				uint64_t q = ntc::mix64(nn + t + 1);
				uint32_t start = (uint32_t)(q % L), rl = 1 + (uint32_t)((q >> 32) % 20);
				for (uint32_t a = start; a < start + rl && a < L; a++)
					o[a] = 'N';
			}
		}
	}
	return NTC_OK;
}

int ntc_gen_packed(uint64_t seed, uint64_t first, uint64_t n, uint32_t L, int mode, uint64_t U, uint32_t stride_words,
    uint32_t* words)
{
	if (!words || mode < 0 || mode > 1 || stride_words < 1 + (L + 15) / 16)
		return set_err(NTC_EINVAL, "ntc_gen_packed: bad argument");
	std::vector<uint8_t> codes(L + 16);
	for (uint64_t r = 0; r < n; r++) {
		ntc::gen_codes(seed, first + r, L, mode, U, codes.data());
		uint32_t* o = words + r * (uint64_t)stride_words;
		memset(o, 0, stride_words * sizeof(uint32_t));
		o[0] = L;
		for (uint32_t j = 0; j < L; j++)
			o[1 + (j >> 4)] |= (uint32_t)codes[j] << (2 * (j & 15));
	}
	return NTC_OK;
}

// compEst, ntcard.cpp:237-275, in the reference's operation order (plain double, no FMA contraction:
// this file is compiled with -ffp-contract=off).  The recurrence stops at covMax: f[i] depends only
// on f[j], j < i (ntcard.cpp:266-272), so rows 1..covMax equal the full 65535-step run bit for bit.
int ntc_estimate(const uint32_t* p_hist, const uint16_t* t_Counter, unsigned rBits, unsigned sBits, unsigned covMax, double* F0,
    double* f)
{
	if ((!p_hist) == (!t_Counter) || !F0 || !f || rBits < 1 || rBits > 40 || sBits + rBits > 63)
		return set_err(NTC_EINVAL, "ntc_estimate: bad argument");
	if (covMax > 65535)
		covMax = 65535; // ntcard.cpp:342-343
	const size_t nSamp = NTC_NSAMP;
	const uint64_t rBuck = (uint64_t)1 << rBits;
	std::vector<unsigned> p(nSamp * 65536, 0u);
	if (t_Counter) {
		for (size_t i = 0; i < nSamp; i++)
			for (uint64_t j = 0; j < rBuck; j++)
				++p[i * 65536 + t_Counter[i * rBuck + j]]; // ntcard.cpp:245-247
	} else {
		for (size_t i = 0; i < nSamp * 65536; i++)
			p[i] = p_hist[i];
	}
	std::vector<double> pMean(65536, 0.0);
	for (size_t i = 0; i < 65536; i++) {
		for (size_t j = 0; j < nSamp; j++)
			pMean[i] += p[j * 65536 + i];
		pMean[i] /= 1.0 * nSamp; // ntcard.cpp:252-256
	}
	const double F0Mean = (ssize_t)((rBits * log(2) - log(pMean[0])) * 1.0 * ((uint64_t)1 << (sBits + rBits))); // :258-259
	*F0 = F0Mean;
	for (unsigned i = 0; i <= covMax; i++)
		f[i] = 0;
	if (pMean[0] * (log(pMean[0]) - rBits * log(2)) == 0) // :262-264
		return NTC_OK;
	if (covMax >= 1)
		f[1] = -1.0 * pMean[1] / (pMean[0] * (log(pMean[0]) - rBits * log(2))); // :265
	for (size_t i = 2; i <= covMax; i++) {
		double sum = 0.0;
		for (size_t j = 1; j < i; j++)
			sum += j * pMean[i - j] * f[j];
		f[i] = -1.0 * pMean[i] / (pMean[0] * (log(pMean[0]) - rBits * log(2))) - sum / (i * pMean[0]); // :266-272
	}
	for (size_t i = 1; i <= covMax; i++)
		f[i] = labs((ssize_t)(f[i] * F0Mean)); // :273-274
	return NTC_OK;
}

// What ntc_submit checks of a ragged batch, as one branch-free pass the compiler vectorises (a batch has millions of records
// and the loop runs on the submitting thread): offsets ascend strictly -- every record has at least its length word -- and end
// inside the batch.
int ntc_check_offsets(const uint32_t* off, size_t n_rec, size_t n_words, uint32_t* max_rec_words)
{
	if (!off)
		return set_err(NTC_EINVAL, "ntc_check_offsets: bad argument");
	uint32_t mx = 0, bad = 0;
	for (size_t i = 0; i < n_rec; i++) {
		const uint32_t a = off[i], b = off[i + 1];
		bad |= (uint32_t)(b <= a);
		const uint32_t d = b - a;
		mx = d > mx ? d : mx;
	}
	if (bad)
		for (size_t i = 0; i < n_rec; i++)
			if (off[i + 1] <= off[i])
				return set_err(NTC_EINVAL, "ntc_submit: record %zu has no length word", i);
	if (off[n_rec] > n_words)
		return set_err(NTC_EINVAL, "ntc_submit: off[n_rec] exceeds n_words");
	if (max_rec_words)
		*max_rec_words = mx;
	return NTC_OK;
}

// The estimate nthll prints, nthll.cpp:243-254, in the reference's operation order: alpha = 1.4426 / (1 + 1.079 / nBuck),
// halved for canonical k-mers (:245; the reference's opt::canon is always true), harmonic mean of 2^register.
int ntc_hll_estimate(const uint8_t* regs, unsigned nBits, int canon, double* est)
{
	if (!regs || !est || nBits < 1 || nBits > 30)
		return set_err(NTC_EINVAL, "ntc_hll_estimate: bad argument");
	const unsigned nBuck = ((unsigned)1) << nBits;
	double pEst = 0.0, zEst = 0.0, eEst = 0.0, alpha = 0.0;
	alpha = 1.4426 / (1 + 1.079 / nBuck);
	if (canon)
		alpha /= 2;
	for (unsigned j = 0; j < nBuck; j++) {
		if (regs[j] > 63)
			return set_err(NTC_EINVAL, "ntc_hll_estimate: register %u holds %u (> 63)", j, (unsigned)regs[j]);
		pEst += 1.0 / ((uint64_t)1 << regs[j]);
	}
	zEst = 1.0 / pEst;
	eEst = alpha * nBuck * nBuck * zEst;
	*est = eEst;
	return NTC_OK;
}

} // extern "C"
