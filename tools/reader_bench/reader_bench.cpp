// reader_bench.cpp -- host ingest throughput of the CLI without a device (see stub_device.cpp):
//   reader_bench K THREADS FILE...      -> GB/s of input text, records and words produced, order-independent digest of the records
// THREADS > number of files: the spare threads parse pieces of the same file (read_file_parallel), as bin/ntcard -t does
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "reader.h"

extern "C" uint64_t stub_words();
extern "C" uint64_t stub_recs();
extern "C" uint64_t stub_batches();
extern "C" uint64_t stub_digest();

int main(int argc, char** argv)
{
	if (argc < 4) {
		fprintf(stderr, "usage: reader_bench K THREADS FILE...\n");
		return 2;
	}
	const unsigned k = (unsigned)atoi(argv[1]);
	unsigned nthreads = (unsigned)atoi(argv[2]);
	std::vector<std::string> files(argv + 3, argv + argc);
	uint64_t bytes = 0;
	for (auto& f : files) {
		struct stat st;
		if (stat(f.c_str(), &st) == 0)
			bytes += (uint64_t)st.st_size;
	}
	std::atomic<size_t> next(0);
	std::mutex mu;
	const auto t0 = std::chrono::steady_clock::now();
	const unsigned per_file = nthreads > files.size() ? nthreads / (unsigned)files.size() : 1;
	if (nthreads > files.size())
		nthreads = (unsigned)files.size();
	const std::function<ntc_ctx*()> no_ctx = []() { return (ntc_ctx*)nullptr; };
	auto worker = [&]() {
		ntcb::BatchSubmitter sub(nullptr, k, &mu);
		for (;;) {
			size_t i = next.fetch_add(1);
			if (i >= files.size())
				break;
			bool handled = false;
			if (per_file > 1)
				ntcb::read_file_parallel(files[i], no_ctx, k, &mu, per_file, &handled);
			if (!handled && !ntcb::read_file(files[i], sub))
				fprintf(stderr, "cannot read %s\n", files[i].c_str());
		}
		sub.flush();
		sub.finish();
	};
	std::vector<std::thread> th;
	for (unsigned t = 0; t < nthreads; t++)
		th.emplace_back(worker);
	for (auto& t : th)
		t.join();
	const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	printf("%.3f s, %.2f GB/s of text (%u x %u threads, %zu files, %.2f GB); %llu records, %llu batches, digest %016llx\n", s, bytes / s / 1e9, nthreads,
	    per_file, files.size(), bytes / 1e9, (unsigned long long)stub_recs(), (unsigned long long)stub_batches(), (unsigned long long)stub_digest());
	return 0;
}
