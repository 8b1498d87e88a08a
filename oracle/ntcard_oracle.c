/*
 * oracle/ntcard_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the one ntCard hot path this repository rebuilds
 * for sm_100a: canonical ntHash over every valid k-mer of a sequence, the
 * two-table sampling test, the uint16 sketch increment, and the estimator.
 * Every function cites the reference file:line it follows (paths relative to
 * the upstream tree, bcgsc/ntCard v1.2.2).
 *
 * Nothing under ntcard_b200/ (the product) may call, link or load this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker or the timed CPU arm.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - the ntHash known-answer vectors of vendor/ntHash/unittest/UnitTests.cpp:39-93,
 *   - tests/golden/ fixtures generated from the reference's own code compiled
 *     here (oracle/_ref, see oracle/Makefile and oracle/make_golden.py),
 *   - and, when oracle/_ref/libntcard_ref.so is present, the reference itself
 *     on random inputs.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- seeds: vendor/ntHash/nthash.hpp:25-64 --------------------------------- */
#define ORC_SEED_A 0x3c8bfbb395c60474ULL
#define ORC_SEED_C 0x3193c18562a02b4cULL
#define ORC_SEED_G 0x20323ed082572324ULL
#define ORC_SEED_T 0x295549f54be24456ULL
#define ORC_SEED_N 0x0000000000000000ULL

/* seedTab[256], nthash.hpp:31-64.  Row 0..7 holds the complement seeds reached
 * through `c & cpOff` (cpOff = 0x07, nthash.hpp:16): 'A'&7=1 -> T, 'C'&7=3 -> G,
 * 'G'&7=7 -> C, 'T'&7=4 -> A, 'U'&7=5 -> A.  Lower case has the same low bits. */
uint64_t orc_seed(unsigned char c)
{
	switch (c) {
	case 1: return ORC_SEED_T;
	case 3: return ORC_SEED_G;
	case 4: return ORC_SEED_A;
	case 5: return ORC_SEED_A;
	case 7: return ORC_SEED_C;
	case 'A': case 'a': return ORC_SEED_A;
	case 'C': case 'c': return ORC_SEED_C;
	case 'G': case 'g': return ORC_SEED_G;
	case 'T': case 't': case 'U': case 'u': return ORC_SEED_T;
	default: return ORC_SEED_N;
	}
}

/* nthash.hpp:186-217: rol1 + swapbits033 rotate the upper 31 bits and the lower
 * 33 bits as two independent rings by one position to the left; ror1 +
 * swapbits3263 is the inverse. */
uint64_t orc_srol(uint64_t v)
{
	uint64_t r = (v << 1) | (v >> 63);           /* rol1, :186-188 */
	uint64_t x = (r ^ (r >> 33)) & 1;            /* swapbits033, :208-211 */
	return r ^ (x | (x << 33));
}

uint64_t orc_sror(uint64_t v)
{
	uint64_t r = (v >> 1) | (v << 63);           /* ror1, :191-193 */
	uint64_t x = ((r >> 32) ^ (r >> 63)) & 1;    /* swapbits3263, :214-217 */
	return r ^ ((x << 32) | (x << 63));
}

/* srol applied n times.  Equals msTab31l[c][n%31] | msTab33r[c][n%33] of
 * nthash.hpp:66-183 when v = seedTab[c] (checked in tests against _ref). */
uint64_t orc_srol_n(uint64_t v, unsigned n)
{
	uint64_t hi = v >> 33, lo = v & 0x1FFFFFFFFULL;
	unsigned a = n % 31, b = n % 33;
	if (a) hi = ((hi << a) | (hi >> (31 - a))) & 0x7FFFFFFFULL;   /* rol31, :196-199 */
	if (b) lo = ((lo << b) | (lo >> (33 - b))) & 0x1FFFFFFFFULL;  /* rol33, :202-205 */
	return (hi << 33) | lo;
}

/* From-scratch forward / reverse-complement hash: nthash.hpp:220-239. */
uint64_t orc_ntf64(const char* kmer, unsigned k)
{
	uint64_t h = 0;
	for (unsigned i = 0; i < k; i++)
		h = orc_srol(h) ^ orc_seed((unsigned char)kmer[i]);
	return h;
}

uint64_t orc_ntr64(const char* kmer, unsigned k)
{
	uint64_t h = 0;
	for (unsigned i = 0; i < k; i++)
		h = orc_srol(h) ^ orc_seed((unsigned char)kmer[k - 1 - i] & 7);
	return h;
}

/* Rolling updates: nthash.hpp:242-257. */
uint64_t orc_ntf64_roll(uint64_t fh, unsigned k, unsigned char out, unsigned char in)
{
	return orc_srol(fh) ^ orc_seed(in) ^ orc_srol_n(orc_seed(out), k);
}

uint64_t orc_ntr64_roll(uint64_t rh, unsigned k, unsigned char out, unsigned char in)
{
	uint64_t h = rh ^ orc_srol_n(orc_seed(in & 7), k);
	h ^= orc_seed(out & 7);
	return orc_sror(h);
}

/* NTMC64 base form with validity scan, m = 1: nthash.hpp:467-492.  Returns 0 and
 * the index of the RIGHTMOST invalid character in *locN, else 1 and fh/rh. */
static int orc_ntmc64_base(const char* kmer, unsigned k, uint64_t* fh, uint64_t* rh, unsigned* locN)
{
	uint64_t f = 0, r = 0;
	*locN = 0;
	for (int i = (int)k - 1; i >= 0; i--) {
		if (orc_seed((unsigned char)kmer[i]) == ORC_SEED_N) {
			*locN = (unsigned)i;
			return 0;
		}
		f = orc_srol(f) ^ orc_seed((unsigned char)kmer[k - 1 - i]);
		r = orc_srol(r) ^ orc_seed((unsigned char)kmer[i] & 7);
	}
	*fh = f;
	*rh = r;
	return 1;
}

/* ntHashIterator restated as a state machine: ntHashIterator.hpp:59-86. */
typedef struct {
	const char* seq;
	size_t len;
	unsigned k;
	size_t pos; /* SIZE_MAX == end */
	uint64_t fh, rh;
} orc_iter;

static void orc_iter_init(orc_iter* it) /* ntHashIterator.hpp:59-70 */
{
	if (it->k > it->len) {
		it->pos = SIZE_MAX;
		return;
	}
	unsigned locN = 0;
	while (it->pos < it->len - it->k + 1 &&
	       !orc_ntmc64_base(it->seq + it->pos, it->k, &it->fh, &it->rh, &locN))
		it->pos += locN + 1;
	if (it->pos >= it->len - it->k + 1)
		it->pos = SIZE_MAX;
}

static void orc_iter_next(orc_iter* it) /* ntHashIterator.hpp:73-86 */
{
	++it->pos;
	if (it->pos >= it->len - it->k + 1) {
		it->pos = SIZE_MAX;
		return;
	}
	unsigned char in = (unsigned char)it->seq[it->pos + it->k - 1];
	if (orc_seed(in) == ORC_SEED_N) {
		it->pos += it->k;
		orc_iter_init(it);
	} else {
		unsigned char out = (unsigned char)it->seq[it->pos - 1];
		it->fh = orc_ntf64_roll(it->fh, it->k, out, in);  /* NTC64 roll, nthash.hpp:275-279 */
		it->rh = orc_ntr64_roll(it->rh, it->k, out, in);
	}
}

static inline uint64_t orc_iter_hash(const orc_iter* it) /* min(fh, rh), nthash.hpp:278 */
{
	return it->rh < it->fh ? it->rh : it->fh;
}

/* All canonical hashes of a sequence in iterator order.  Returns the number of
 * valid k-mers; fills at most cap entries of out_h / out_pos (either may be NULL). */
size_t orc_hash_seq(const char* seq, size_t len, unsigned k, uint64_t* out_h, uint32_t* out_pos, size_t cap)
{
	orc_iter it = { seq, len, k, 0, 0, 0 };
	size_t n = 0;
	if (k == 0)
		return 0;
	orc_iter_init(&it);
	while (it.pos != SIZE_MAX) {
		if (n < cap) {
			if (out_h) out_h[n] = orc_iter_hash(&it);
			if (out_pos) out_pos[n] = (uint32_t)it.pos;
		}
		++n;
		orc_iter_next(&it);
	}
	return n;
}

/* Forward and reverse hashes of one k-mer (for unit tests). */
void orc_kmer_hashes(const char* kmer, unsigned k, uint64_t* fh, uint64_t* rh)
{
	*fh = orc_ntf64(kmer, k);
	*rh = orc_ntr64(kmer, k);
}

/* ntComp: ntcard.cpp:132-145.  Returns the table (0/1) or 2 when not sampled. */
static inline unsigned orc_ntcomp(uint64_t h, uint16_t* t, unsigned rBits, unsigned sBits, int atomic)
{
	const uint64_t rBuck = (uint64_t)1 << rBits;
	const uint64_t sMask = ((uint64_t)1 << (sBits - 1)) - 1; /* ntcard.cpp:438 */
	unsigned ind = 2;
	if (h >> (63 - sBits) == 1)
		ind = 0;
	if (h >> (64 - sBits) == sMask)
		ind = 1;
	if (ind < 2) {
		uint16_t* c = &t[ind * rBuck + (h & (rBuck - 1))];
		if (atomic) {
#pragma omp atomic
			++*c;
		} else
			++*c;
	}
	return ind;
}

unsigned orc_sample_table(uint64_t h, unsigned sBits)
{
	const uint64_t sMask = ((uint64_t)1 << (sBits - 1)) - 1;
	unsigned ind = 2;
	if (h >> (63 - sBits) == 1) ind = 0;
	if (h >> (64 - sBits) == sMask) ind = 1;
	return ind;
}

/* ntRead: ntcard.cpp:147-158.  t_Counter layout [kIdx][table][bucket], uint16. */
static void orc_ntread_impl(const char* seq, size_t len, const unsigned* kList, unsigned nK,
    unsigned rBits, unsigned sBits, uint16_t* t, uint64_t* totKmer, int atomic)
{
	const uint64_t rBuck = (uint64_t)1 << rBits;
	for (unsigned ki = 0; ki < nK; ki++) {
		orc_iter it = { seq, len, kList[ki], 0, 0, 0 };
		orc_iter_init(&it);
		while (it.pos != SIZE_MAX) {
			orc_ntcomp(orc_iter_hash(&it), t + (uint64_t)ki * 2 * rBuck, rBits, sBits, atomic);
			orc_iter_next(&it);
			++totKmer[ki];
		}
	}
}

void orc_ntread(const char* seq, size_t len, const unsigned* kList, unsigned nK,
    unsigned rBits, unsigned sBits, uint16_t* t, uint64_t* totKmer)
{
	orc_ntread_impl(seq, len, kList, nK, rBits, sBits, t, totKmer, 0);
}

/* A batch of sequences: seqs is one byte array, sequence i = [off[i], off[i+1]).
 * With nthreads > 1 the reads are split over OpenMP threads on the shared
 * sketch, the way ntcard.cpp:445-467 does per file (atomic ++, per-thread F1). */
void orc_ntread_batch(const char* seqs, const uint64_t* off, size_t n, const unsigned* kList,
    unsigned nK, unsigned rBits, unsigned sBits, uint16_t* t, uint64_t* totKmer, int nthreads)
{
	if (nthreads <= 1) {
		for (size_t i = 0; i < n; i++)
			orc_ntread_impl(seqs + off[i], off[i + 1] - off[i], kList, nK, rBits, sBits, t, totKmer, 0);
		return;
	}
#pragma omp parallel num_threads(nthreads)
	{
		uint64_t loc[64];
		for (unsigned ki = 0; ki < nK && ki < 64; ki++) loc[ki] = 0;
#pragma omp for schedule(dynamic, 4096)
		for (size_t i = 0; i < n; i++)
			orc_ntread_impl(seqs + off[i], off[i + 1] - off[i], kList, nK, rBits, sBits, t, loc, 1);
		for (unsigned ki = 0; ki < nK && ki < 64; ki++) {
#pragma omp atomic
			totKmer[ki] += loc[ki];
		}
	}
}

/* ---- gap seeds (-g N): stRead, ntcard.cpp:160-171, with the spaced seed 1..10..01..1 of ntcard.cpp:407-413 ----
 * stHashIterator (stHashIterator.hpp:60-94) walks the sequence like ntHashIterator; per window NTMSM64
 * (nthash.hpp:620-678, m = m2 = 1) takes the plain fh/rh and XORs the contribution of every '0' position i
 * of the seed back out:  fs = fh ^ srol^(k-1-i)(seed[c_i]),  rs = rh ^ srol^i(seed[comp c_i]);  h = min(fs, rs).
 * The seed string has (k-gap)/2 ones, gap zeros, (k-gap)/2 ones (k and gap of equal parity, ntcard.cpp:382). */
static inline uint64_t orc_st_hash(const char* w, unsigned k, unsigned gap, uint64_t fh, uint64_t rh)
{
	const unsigned a = (k - gap) / 2;
	uint64_t fs = fh, rs = rh;
	for (unsigned i = a; i < a + gap; i++) {
		fs ^= orc_srol_n(orc_seed((unsigned char)w[i]), k - 1 - i);
		rs ^= orc_srol_n(orc_seed((unsigned char)w[i] & 7), i);
	}
	return rs < fs ? rs : fs;
}

static void orc_stread_impl(const char* seq, size_t len, unsigned k, unsigned gap, unsigned rBits, unsigned sBits, uint16_t* t,
    uint64_t* totKmer, int atomic)
{
	orc_iter it = { seq, len, k, 0, 0, 0 }; /* the position logic of stHashIterator::init/next equals ntHashIterator's */
	orc_iter_init(&it);
	while (it.pos != SIZE_MAX) {
		orc_ntcomp(orc_st_hash(seq + it.pos, k, gap, it.fh, it.rh), t, rBits, sBits, atomic);
		orc_iter_next(&it);
		++totKmer[0];
	}
}

/* All gap-seed hashes of a sequence in iterator order (unit tests). */
size_t orc_st_hash_seq(const char* seq, size_t len, unsigned k, unsigned gap, uint64_t* out_h, size_t cap)
{
	orc_iter it = { seq, len, k, 0, 0, 0 };
	size_t n = 0;
	orc_iter_init(&it);
	while (it.pos != SIZE_MAX) {
		if (n < cap)
			out_h[n] = orc_st_hash(seq + it.pos, k, gap, it.fh, it.rh);
		++n;
		orc_iter_next(&it);
	}
	return n;
}

void orc_stread_batch(const char* seqs, const uint64_t* off, size_t n, unsigned k, unsigned gap, unsigned rBits, unsigned sBits,
    uint16_t* t, uint64_t* totKmer, int nthreads)
{
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
	{
		uint64_t loc = 0;
#pragma omp for schedule(dynamic, 4096)
		for (size_t i = 0; i < n; i++)
			orc_stread_impl(seqs + off[i], off[i + 1] - off[i], k, gap, rBits, sBits, t, &loc, 1);
#pragma omp atomic
		totKmer[0] += loc;
	}
}

/* ---- nthll: the HyperLogLog F0 estimator that ships beside ntcard (nthll.cpp) -----------------------------------
 * ntComp (nthll.cpp:92-97): with nBuck = 2^nBits registers, a canonical hash h whose bits above the low nBits are not
 * all zero updates register h & (nBuck-1) with run0 = clz(h & ~(nBuck-1)) if that is larger.  ntRead (nthll.cpp:99-104)
 * walks the sequence with the same ntHashIterator as ntcard; the readers skip sequences shorter than k
 * (nthll.cpp:112,129,146).  Threads keep private registers merged by max (nthll.cpp:213-241). */
static void orc_hll_read(const char* seq, size_t len, unsigned k, unsigned nBits, uint8_t* m)
{
	const uint64_t nBuck = (uint64_t)1 << nBits;
	orc_iter it = { seq, len, k, 0, 0, 0 };
	if (len < k)
		return;
	orc_iter_init(&it);
	while (it.pos != SIZE_MAX) {
		const uint64_t h = orc_iter_hash(&it), top = h & ~(nBuck - 1);
		if (top) {
			const uint8_t run0 = (uint8_t)__builtin_clzll(top);
			if (run0 > m[h & (nBuck - 1)])
				m[h & (nBuck - 1)] = run0;
		}
		orc_iter_next(&it);
	}
}

void orc_hll_batch(const char* seqs, const uint64_t* off, size_t n, unsigned k, unsigned nBits, uint8_t* regs, int nthreads)
{
	const uint64_t nBuck = (uint64_t)1 << nBits;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
	{
		uint8_t* m = (uint8_t*)calloc(nBuck, 1);
#pragma omp for schedule(dynamic, 1024)
		for (size_t i = 0; i < n; i++)
			orc_hll_read(seqs + off[i], off[i + 1] - off[i], k, nBits, m);
#pragma omp critical(orc_hll_merge)
		for (uint64_t j = 0; j < nBuck; j++)
			if (regs[j] < m[j])
				regs[j] = m[j];
		free(m);
	}
}

/* The estimate main() prints (nthll.cpp:243-254), canonical k-mers (alpha / 2, nthll.cpp:245). */
double orc_hll_estimate(const uint8_t* regs, unsigned nBits)
{
	const unsigned nBuck = 1u << nBits;
	double pEst = 0.0, zEst, eEst, alpha;
	alpha = 1.4426 / (1 + 1.079 / nBuck);
	alpha /= 2;
	for (unsigned j = 0; j < nBuck; j++)
		pEst += 1.0 / ((uint64_t)1 << regs[j]);
	zEst = 1.0 / pEst;
	eEst = alpha * nBuck * nBuck * zEst;
	return eEst;
}

/* compEst: ntcard.cpp:237-275.  f must hold 65536 doubles.  imax (2..65535)
 * truncates the recurrence; f[i] for i <= imax is identical to the full run
 * because f[i] depends only on f[j], j < i (ntcard.cpp:266-272).
 * p_hist (2 x 65536 uint32, counter-value histogram, ntcard.cpp:245-247) may be
 * passed instead of t when t == NULL. */
void orc_compest(const uint16_t* t, const uint32_t* p_hist, unsigned rBits, unsigned sBits,
    unsigned imax, double* F0Mean, double* fMean)
{
	const size_t nSamp = 2;
	const uint64_t rBuck = (uint64_t)1 << rBits;
	unsigned* p = (unsigned*)calloc(nSamp * 65536, sizeof(unsigned));
	double* pMean = (double*)calloc(65536, sizeof(double));
	if (t) {
		for (size_t i = 0; i < nSamp; i++)
			for (uint64_t j = 0; j < rBuck; j++)
				++p[i * 65536 + t[i * rBuck + j]];
	} else {
		for (size_t i = 0; i < nSamp * 65536; i++)
			p[i] = p_hist[i];
	}
	for (size_t i = 0; i < 65536; i++) {
		for (size_t j = 0; j < nSamp; j++)
			pMean[i] += p[j * 65536 + i];
		pMean[i] /= 1.0 * nSamp;
	}
	*F0Mean = (ssize_t)((rBits * log(2) - log(pMean[0])) * 1.0 * ((uint64_t)1 << (sBits + rBits)));
	for (size_t i = 0; i < 65536; i++)
		fMean[i] = 0;
	if (pMean[0] * (log(pMean[0]) - rBits * log(2)) == 0)
		goto done;
	if (imax > 65535) imax = 65535;
	fMean[1] = -1.0 * pMean[1] / (pMean[0] * (log(pMean[0]) - rBits * log(2)));
	for (size_t i = 2; i <= imax; i++) {
		double sum = 0.0;
		for (size_t j = 1; j < i; j++)
			sum += j * pMean[i - j] * fMean[j];
		fMean[i] = -1.0 * pMean[i] / (pMean[0] * (log(pMean[0]) - rBits * log(2))) - sum / (i * pMean[0]);
	}
	for (size_t i = 1; i <= imax; i++)
		fMean[i] = labs((ssize_t)(fMean[i] * *F0Mean));
done:
	free(p);
	free(pMean);
}

/* outDefault body for one k: ntcard.cpp:291-294.  Returns 0 on success. */
int orc_write_hist(const char* path, uint64_t F1, double F0Mean, const double* fMean, unsigned covMax)
{
	FILE* fp = fopen(path, "w");
	if (!fp) return -1;
	fprintf(fp, "F1\t%llu\n", (unsigned long long)F1);
	fprintf(fp, "F0\t%llu\n", (unsigned long long)(uint64_t)F0Mean);
	for (unsigned i = 1; i <= covMax; i++)
		fprintf(fp, "%u\t%llu\n", i, (unsigned long long)(uint64_t)fMean[i]);
	fclose(fp);
	return 0;
}

/* ---- deterministic synthetic reads (ours, not the reference's) ---- */
uint64_t orc_mix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ULL;
	uint64_t z = x;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

static void orc_gen_plain(uint64_t S, uint64_t i, unsigned L, char* out)
{
	static const char B[4] = { 'A', 'C', 'G', 'T' };
	for (unsigned w = 0; w * 32 < L; w++) {
		uint64_t v = orc_mix64((S << 48) ^ (i << 12) ^ w);
		for (unsigned j = 0; j < 32 && w * 32 + j < L; j++)
			out[w * 32 + j] = B[(v >> (2 * j)) & 3];
	}
}

/* mode 0 uniform; mode 1 repeat (read i = gen(i mod U), reverse-complemented
 * when (i / U) is odd); mode 2 = uniform + N runs. */
void orc_gen_read(uint64_t S, uint64_t i, unsigned L, int mode, uint64_t U, char* out)
{
	if (mode == 1 && U) {
		orc_gen_plain(S, i % U, L, out);
		if ((i / U) & 1) {
			for (unsigned a = 0, b = L - 1; a < b; a++, b--) {
				char t = out[a]; out[a] = out[b]; out[b] = t;
			}
			for (unsigned a = 0; a < L; a++) {
				switch (out[a]) {
				case 'A': out[a] = 'T'; break;
				case 'C': out[a] = 'G'; break;
				case 'G': out[a] = 'C'; break;
				default: out[a] = 'A'; break;
				}
			}
		}
		return;
	}
	orc_gen_plain(S, i, L, out);
	if (mode == 2) {
		uint64_t nn = orc_mix64((S << 48) ^ (i << 12) ^ 0xFFF);
		unsigned runs = (unsigned)(nn & 3);
		for (unsigned t = 0; t < runs; t++) {
			uint64_t q = orc_mix64(nn + t + 1);
			unsigned start = (unsigned)(q % L), rl = 1 + (unsigned)((q >> 32) % 20);
			for (unsigned a = start; a < start + rl && a < L; a++)
				out[a] = 'N';
		}
	}
}

void orc_gen_reads(uint64_t S, uint64_t first, uint64_t n, unsigned L, int mode, uint64_t U, char* out)
{
#pragma omp parallel for schedule(static)
	for (uint64_t i = 0; i < n; i++)
		orc_gen_read(S, first + i, L, mode, U, out + i * (uint64_t)L);
}

/* FNV-1a-64 digest over (idx,count) of nonzero buckets of one table, in index
 * order (SURVEY.md 8c fixture FX1). */
uint64_t orc_table_digest(const uint16_t* t, uint64_t n, uint64_t* nnz, uint64_t* sum, unsigned* maxv, uint64_t* first)
{
	uint64_t h = 0xcbf29ce484222325ULL, nz = 0, s = 0, f = (uint64_t)-1;
	unsigned m = 0;
	for (uint64_t i = 0; i < n; i++) {
		if (t[i]) {
			h = (h ^ i) * 0x100000001b3ULL;
			h = (h ^ t[i]) * 0x100000001b3ULL;
			nz++;
			s += t[i];
			if (t[i] > m) m = t[i];
			if (f == (uint64_t)-1) f = i;
		}
	}
	if (nnz) *nnz = nz;
	if (sum) *sum = s;
	if (maxv) *maxv = m;
	if (first) *first = f;
	return h;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
