/*
 * oracle/nthll_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin extern "C" shims around the UNMODIFIED reference nthll.cpp (the HyperLogLog F0 estimator that ships next to
 * ntcard), compiled from where it lies (-I$(REF) -I$(REF)/Common); main() is renamed so that ref_hll_* below call the
 * reference's own ntRead / ntComp (nthll.cpp:92-104).  A separate shared object from ref_harness.cpp because both
 * reference translation units define ntRead / ntComp / getEfq.  Output goes to oracle/_ref/ only (git-ignored).
 */
#define main nthll_reference_main
#include "nthll.cpp" /* resolved through -I<reference root> */
#undef main

#ifdef REF_HARNESS_STUB_UNCOMPRESS
bool
uncompress_init() /* see ref_harness.cpp: the shim must not interpose fopen inside a Python process */
{
	return false;
}
#endif

extern "C" {

void
ref_hll_set(unsigned k, unsigned nBits) /* the globals main() sets: nthll.cpp:43-53, 205-206 */
{
	opt::kmLen = k;
	opt::nBits = nBits;
	opt::nBuck = ((unsigned)1) << nBits;
	opt::canon = true;
}

/* The reference's ntRead over a batch of sequences held in RAM, with the readers' length test (nthll.cpp:112, 129, 146);
 * per-thread registers merged by max, as main() does (nthll.cpp:213-241). */
void
ref_hll_batch(const char* seqs, const uint64_t* off, size_t n, uint8_t* regs, int nthreads)
{
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
	{
		std::vector<uint8_t> m(opt::nBuck, 0);
		std::string s;
#pragma omp for schedule(dynamic, 1024)
		for (size_t i = 0; i < n; i++) {
			s.assign(seqs + off[i], off[i + 1] - off[i]);
			if (s.length() >= opt::kmLen)
				ntRead(s, m.data());
		}
#pragma omp critical(vmrg)
		for (unsigned j = 0; j < opt::nBuck; j++)
			if (regs[j] < m[j])
				regs[j] = m[j];
	}
}

} /* extern "C" */
