// tools/bitslice_selftest.cu -- CPU self-test of bitslice_core.cuh (no GPU needed):
//   nvcc -O2 -std=c++17 --expt-relaxed-constexpr -o /tmp/bs_selftest tools/bitslice_selftest.cu && /tmp/bs_selftest
// Checks (1) transpose32, (2) that the bit-sliced filter marks exactly the k-mers ntComp samples.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ntcard_b200/csrc/bitslice_core.cuh"

using namespace ntc;

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t rnd()
{
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return rng_state;
}

template <int KM, int S, int TQ> struct Body {
	static bool run(bs::State& st, const std::vector<uint32_t>& lo, const std::vector<uint32_t>& hi, int q0, int n, int k,
	    std::vector<uint32_t>& masks)
	{
		const int q = q0 + TQ;
		if (q >= n)
			return false;
		const uint32_t olo = q >= k ? lo[q - k] : 0, ohi = q >= k ? hi[q - k] : 0;
		bs::step<KM, TQ>(st, lo[q], hi[q], olo, ohi);
		if (q >= k - 1)
			masks[q - k + 1] = bs::sampled_mask<TQ, S>(st);
		return Body<KM, S, TQ + 1>::run(st, lo, hi, q0, n, k, masks);
	}
};
template <int KM, int S> struct Body<KM, S, 31> {
	static bool run(bs::State&, const std::vector<uint32_t>&, const std::vector<uint32_t>&, int, int, int, std::vector<uint32_t>&) { return true; }
};

template <int KM, int S> int check(int k, int n)
{
	if (k % 31 != KM) { printf("bad KM\n"); return 1; }
	std::vector<std::vector<uint8_t>> reads(32, std::vector<uint8_t>(n));
	for (auto& r : reads)
		for (auto& c : r)
			c = rnd() & 3;
	// low-complexity slots to exercise constants: all-A, all-T, alternating
	for (int i = 0; i < n; i++) { reads[0][i] = 0; reads[1][i] = 3; reads[2][i] = i & 1 ? 1 : 2; }
	std::vector<uint32_t> lo(n, 0), hi(n, 0);
	for (int s = 0; s < 32; s++)
		for (int i = 0; i < n; i++) {
			lo[i] |= (uint32_t)(reads[s][i] & 1) << s;
			hi[i] |= (uint32_t)(reads[s][i] >> 1) << s;
		}
	bs::State st;
	bs::init_state(k, st.F, st.R);
	std::vector<uint32_t> masks(n, 0);
	for (int q0 = 0; q0 < n; q0 += 31)
		if (!Body<KM, S, 0>::run(st, lo, hi, q0, n, k, masks))
			break;
	int bad = 0, nsamp = 0;
	for (int s = 0; s < 32; s++) {
		for (int p = 0; p + k <= n; p++) {
			uint64_t fh = 0, rh = 0;
			for (int i = 0; i < k; i++) fh = srol(fh) ^ seed_of(reads[s][p + i]);
			for (int i = k - 1; i >= 0; i--) rh = srol(rh) ^ seed_of(3 - reads[s][p + i]);
			const uint64_t h = rh < fh ? rh : fh;
			const bool want = sample_table(h, S) < 2;
			const bool got = (masks[p] >> s) & 1;
			nsamp += want;
			if (want != got && bad++ < 5)
				printf("  mismatch k=%d S=%d slot %d p %d want %d got %d h=%016llx\n", k, S, s, p, (int)want, (int)got, (unsigned long long)h);
		}
	}
	printf("k=%3d (KM=%2d) S=%2d n=%d: %d sampled, %d mismatches\n", k, KM, S, n, nsamp, bad);
	return bad;
}

int main()
{
	int bad = 0;
	// transpose
	for (int it = 0; it < 50; it++) {
		uint32_t A[32], B[32];
		for (int i = 0; i < 32; i++) A[i] = B[i] = (uint32_t)rnd();
		bs::transpose32(B);
		for (int o = 0; o < 32; o++)
			for (int s = 0; s < 32; s++)
				if (((B[o] >> s) & 1) != ((A[s] >> o) & 1)) bad++;
	}
	printf("transpose32: %d bad bits\n", bad);
	bad += check<1, 7>(32, 3000);
	bad += check<1, 11>(32, 40000);
	bad += check<2, 7>(64, 3000);
	bad += check<0, 7>(31, 3000);
	bad += check<0, 11>(31, 40000);
	bad += check<12, 7>(12, 3000);
	bad += check<3, 2>(96, 500);
	bad += check<4, 7>(128, 3000);
	bad += check<1, 7>(63, 3000);
	bad += check<5, 3>(5, 1000);
	// every k mod 31 class the library compiles (Makefile BS_KMS = 0..30): k = 31 + KM at s = 7, and the usual assembly k values
#define CLS(KM) bad += check<KM, 7>(31 + KM, 2000);
	CLS(0) CLS(1) CLS(2) CLS(3) CLS(4) CLS(5) CLS(6) CLS(7) CLS(8) CLS(9) CLS(10) CLS(11) CLS(12) CLS(13) CLS(14) CLS(15)
	CLS(16) CLS(17) CLS(18) CLS(19) CLS(20) CLS(21) CLS(22) CLS(23) CLS(24) CLS(25) CLS(26) CLS(27) CLS(28) CLS(29) CLS(30)
#undef CLS
	bad += check<21, 7>(21, 2000);
	bad += check<25, 11>(25, 40000);
	bad += check<10, 7>(41, 2000);
	bad += check<20, 11>(51, 40000);
	bad += check<9, 7>(71, 2000);
	bad += check<8, 7>(101, 2000);
	bad += check<28, 11>(121, 40000);
	bad += check<27, 7>(151, 2000);
	bad += check<14, 7>(200, 2000);
	printf(bad ? "FAILED\n" : "ALL OK\n");
	return bad != 0;
}
