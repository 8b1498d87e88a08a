/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin extern "C" shims around the UNMODIFIED reference, compiled from where it
 * lies (-I$(REF) -I$(REF)/Common; nothing is copied into this repository).  The
 * reference translation unit ntcard.cpp is included whole with its main()
 * renamed, so ref_* below call the reference's own ntRead / ntComp / compEst /
 * ntHashIterator.  Output goes to oracle/_ref/ only (git-ignored).
 *
 * Used to (1) pin oracle/ntcard_oracle.c, (2) generate tests/golden fixtures
 * (oracle/make_golden.py), (3) time the reference's CPU path beside the GPU
 * (bench.py cpu_baseline kind "reference" and --impl reference).
 */
#define main ntcard_reference_main
#include "ntcard.cpp" /* resolved through -I<reference root> */
#undef main

#include <cstring>

/* ntcard.cpp includes Uncompress.h, whose static initialiser calls
 * uncompress_init() (Common/Uncompress.h:4-7).  The real one interposes
 * fopen/open process-wide; the shim library must not do that inside a Python
 * process, so it satisfies the symbol with a no-op.  (The stand-alone CLI built
 * next to it, oracle/_ref/ntcard_ref, links the real Common sources.) */
#ifdef REF_HARNESS_STUB_UNCOMPRESS
bool
uncompress_init()
{
	return false;
}
#endif

extern "C" {

/* Globals the reference reads in ntComp/compEst: ntcard.cpp:52-67, 437-438. */
void
ref_set_opts(unsigned rBits, unsigned sBits, unsigned nK)
{
	opt::rBits = rBits;
	opt::sBits = sBits;
	opt::nK = nK;
	opt::nSamp = 2;
	opt::gap = 0;
	opt::rBuck = ((size_t)1) << opt::rBits;
	opt::sMask = (((size_t)1) << (opt::sBits - 1)) - 1;
}

uint64_t ref_srol(uint64_t v) { return swapbits033(rol1(v)); }
uint64_t ref_sror(uint64_t v) { return swapbits3263(ror1(v)); }
uint64_t ref_seed(unsigned char c) { return seedTab[c]; }
uint64_t ref_mstab(unsigned char c, unsigned k) { return msTab31l[c][k % 31] | msTab33r[c][k % 33]; }

void
ref_kmer_hashes(const char* kmer, unsigned k, uint64_t* fh, uint64_t* rh)
{
	*fh = NTF64(kmer, k);
	*rh = NTR64(kmer, k);
}

size_t
ref_hash_seq(const char* seq, size_t len, unsigned k, uint64_t* out_h, uint32_t* out_pos, size_t cap)
{
	std::string s(seq, len);
	ntHashIterator itr(s, 1, k);
	size_t n = 0;
	while (itr != itr.end()) {
		if (n < cap) {
			if (out_h)
				out_h[n] = (*itr)[0];
			if (out_pos)
				out_pos[n] = (uint32_t)itr.pos();
		}
		++n;
		++itr;
	}
	return n;
}

void
ref_ntread(const char* seq, size_t len, const unsigned* kList, unsigned nK, uint16_t* t, uint64_t* totKmer)
{
	std::vector<unsigned> kl(kList, kList + nK);
	std::string s(seq, len);
	size_t tot[nK];
	for (unsigned i = 0; i < nK; i++)
		tot[i] = 0;
	ntRead(s, kl, t, tot);
	for (unsigned i = 0; i < nK; i++)
		totKmer[i] += tot[i];
}

/* The reference's ntRead over a batch held in RAM, spread over OpenMP threads
 * on the shared sketch (the per-file loop of ntcard.cpp:445-467 applied to
 * reads; SURVEY.md 8d "hot loop only").  Returns seconds spent in the loop. */
double
ref_ntread_batch(const char* seqs, const uint64_t* off, size_t n, const unsigned* kList, unsigned nK,
    uint16_t* t, uint64_t* totKmer, int nthreads)
{
	std::vector<unsigned> kl(kList, kList + nK);
	double t0 = omp_get_wtime();
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
	{
		size_t tot[nK];
		for (unsigned i = 0; i < nK; i++)
			tot[i] = 0;
		std::string s;
#pragma omp for schedule(dynamic, 4096)
		for (size_t i = 0; i < n; i++) {
			s.assign(seqs + off[i], off[i + 1] - off[i]);
			ntRead(s, kl, t, tot);
		}
		for (unsigned i = 0; i < nK; i++) {
#pragma omp atomic
			totKmer[i] += tot[i];
		}
	}
	return omp_get_wtime() - t0;
}

/* Gap-seed mode: the globals main() sets at ntcard.cpp:407-413, then the reference's own stRead (ntcard.cpp:160-171). */
void
ref_set_gap(unsigned k, unsigned gap)
{
	opt::gap = gap;
	opt::seedSet.clear();
	if (gap != 0) {
		std::string g(gap, '0');
		std::string nonGap((k - gap) / 2, '1');
		std::vector<std::string> seedString;
		seedString.push_back(nonGap + g + nonGap);
		opt::seedSet = stHashIterator::parseSeed(seedString);
	}
}

void
ref_stread_batch(const char* seqs, const uint64_t* off, size_t n, unsigned k, uint16_t* t, uint64_t* totKmer, int nthreads)
{
	std::vector<unsigned> kl(1, k);
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
	{
		size_t tot[1] = { 0 };
		std::string s;
#pragma omp for schedule(dynamic, 4096)
		for (size_t i = 0; i < n; i++) {
			s.assign(seqs + off[i], off[i + 1] - off[i]);
			stRead(s, kl, t, tot);
		}
#pragma omp atomic
		totKmer[0] += tot[0];
	}
}

size_t
ref_st_hash_seq(const char* seq, size_t len, unsigned k, uint64_t* out_h, size_t cap)
{
	std::string s(seq, len);
	stHashIterator itr(s, opt::seedSet, 1, 1, k);
	size_t n = 0;
	while (itr != itr.end()) {
		if (n < cap)
			out_h[n] = (*itr)[0];
		++n;
		++itr;
	}
	return n;
}

void
ref_compest(const uint16_t* t, double* F0Mean, double* fMean /* [65536] */)
{
	compEst(t, *F0Mean, fMean);
}

int
ref_max_threads()
{
	return omp_get_max_threads();
}

} /* extern "C" */
