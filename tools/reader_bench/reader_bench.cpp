// reader_bench.cpp -- host ingest throughput of the CLI without a device (see stub_device.cpp):
//   reader_bench K THREADS FILE...      -> GB/s of input text, records and words produced
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
	auto worker = [&]() {
		ntcb::BatchSubmitter sub(nullptr, k, &mu);
		for (;;) {
			size_t i = next.fetch_add(1);
			if (i >= files.size())
				break;
			if (!ntcb::read_file(files[i], sub))
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
	printf("%.3f s, %.2f GB/s of text (%u threads, %zu files, %.2f GB); %llu records, %llu words, %llu batches\n", s, bytes / s / 1e9, nthreads,
	    files.size(), bytes / 1e9, (unsigned long long)stub_recs(), (unsigned long long)stub_words(), (unsigned long long)stub_batches());
	return 0;
}
