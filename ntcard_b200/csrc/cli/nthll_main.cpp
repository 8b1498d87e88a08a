// nthll -- host command line of the B200-native HyperLogLog distinct-k-mer estimator.
//
// Same surface as the reference's nthll main() (nthll.cpp:20-41, 150-256): `nthll [OPTION]... FILES...`, options
// -t -k (+ the undocumented -b register bits, -s, -c, -h of nthll.cpp:55), `@list` arguments, the format sniffer of
// nthll.cpp:73-90, and the one output line on stdout.  The inside differs: reader threads parse and 2-bit pack the
// sequences into pinned double buffers (reader.cpp, shared with ntcard) and the hash -> clz -> max loop
// (ntRead / ntComp, nthll.cpp:92-104) runs on the GPU through the C-ABI (ntc_hll_*, include/ntcard_b200.h); the
// estimate is host arithmetic in the reference's order.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <getopt.h>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "ntcard_b200.h"
#include "reader.h"

#define PROGRAM "nthll"

static const char VERSION_MESSAGE[] = PROGRAM " 1.2.2 (ntcard-b200: sm_100a sketch path)\n";

static const char USAGE_MESSAGE[] =
    "Usage: " PROGRAM " [OPTION]... FILES...\n"
    "Estimates the number of distinct k-mers in FILES(>=1) with HyperLogLog over canonical ntHash.\n"
    "Accepatble file formats: fastq, fasta, sam.\n"
    "\n"
    " Options:\n"
    "\n"
    "  -t, --threads=N	use N parallel reader threads [1] (N>=2 should be used when input files are >=2)\n"
    "  -k, --kmer=N	the length of kmer [64]\n"
    "  -b, --bit=N	log2 of the number of HyperLogLog registers [16]\n"
    "      --gpu=N	CUDA device to use [0]\n"
    "      --help	display this help and exit\n"
    "      --version	output version information and exit\n"
    "\n";

namespace opt {
unsigned nThrd = 1;
unsigned kmLen = 64; // nthll.cpp:45
unsigned nBits = 16; // nthll.cpp:47
unsigned sBits = 22; // nthll.cpp:49, parsed and unused there as well
int gpu = 0;
} // namespace opt

static const char shortopts[] = "t:k:b:s:hc"; // nthll.cpp:55
enum { OPT_HELP = 1, OPT_VERSION, OPT_GPU };
static const struct option longopts[] = { { "threads", required_argument, NULL, 't' }, { "kmer", required_argument, NULL, 'k' },
	{ "bit", required_argument, NULL, 'b' }, { "sit", required_argument, NULL, 's' }, { "gpu", required_argument, NULL, OPT_GPU },
	{ "help", no_argument, NULL, OPT_HELP }, { "version", no_argument, NULL, OPT_VERSION }, { NULL, 0, NULL, 0 } };

static void die_ntc(const char* what)
{
	std::cerr << PROGRAM ": " << what << ": " << ntc_last_error() << "\n";
	exit(EXIT_FAILURE);
}

int main(int argc, char** argv)
{
	bool die = false;
	for (int c; (c = getopt_long(argc, argv, shortopts, longopts, NULL)) != -1;) {
		std::istringstream arg(optarg != NULL ? optarg : "");
		switch (c) {
		case '?': die = true; break;
		case 't': arg >> opt::nThrd; break;
		case 'b': arg >> opt::nBits; break;
		case 's': arg >> opt::sBits; break;
		case 'k': arg >> opt::kmLen; break;
		case 'c': break; // canonical k-mers: always on (nthll.cpp:52, 170-172)
		case 'h': break;
		case OPT_GPU: arg >> opt::gpu; break;
		case OPT_HELP: std::cerr << USAGE_MESSAGE; exit(EXIT_SUCCESS);
		case OPT_VERSION: std::cerr << VERSION_MESSAGE; exit(EXIT_SUCCESS);
		}
		if (optarg != NULL && !arg.eof()) {
			std::cerr << PROGRAM ": invalid option: `-" << (char)c << optarg << "'\n";
			exit(EXIT_FAILURE);
		}
	}
	if (argc - optind < 1) {
		std::cerr << PROGRAM ": missing arguments\n";
		die = true;
	}
	if (die) {
		std::cerr << "Try `" << PROGRAM << " --help' for more information.\n";
		exit(EXIT_FAILURE);
	}
	std::vector<std::string> inFiles; // nthll.cpp:186-197
	for (int i = optind; i < argc; ++i) {
		std::string file(argv[i]);
		if (file[0] == '@') {
			std::string inName;
			std::ifstream inList(file.substr(1, file.length()).c_str());
			while (getline(inList, inName))
				inFiles.push_back(inName);
		} else
			inFiles.push_back(file);
	}

	ntc_ctx* ctx = NULL;
	if (ntc_hll_create(&ctx, opt::kmLen, opt::nBits, opt::gpu, NULL, NULL))
		die_ntc("cannot create the device registers");

	// reader threads: one file at a time per thread, dynamic (nthll.cpp:216 `omp for schedule(dynamic)`); the per-thread
	// registers merged by max (nthll.cpp:232-240) are the one register file on the device
	std::atomic<size_t> next_file(0);
	std::mutex submit_mu;
	unsigned nthreads = opt::nThrd < 1 ? 1 : opt::nThrd;
	if (nthreads > inFiles.size())
		nthreads = (unsigned)inFiles.size();
	auto worker = [&]() {
		ntcb::BatchSubmitter sub(ctx, opt::kmLen, &submit_mu); // sequences shorter than k are dropped: nthll.cpp:112,129,146
		for (;;) {
			size_t i = next_file.fetch_add(1);
			if (i >= inFiles.size())
				break;
			ntcb::read_file(inFiles[inFiles.size() - i - 1], sub, true); // last file first, nthll.cpp:218
		}
		sub.flush();
		sub.finish();
	};
	if (nthreads <= 1) {
		worker();
	} else {
		std::vector<std::thread> th;
		for (unsigned t = 0; t < nthreads; t++)
			th.emplace_back(worker);
		for (auto& t : th)
			t.join();
	}

	std::vector<uint8_t> tVec((size_t)1 << opt::nBits, 0);
	if (ntc_hll_finish(ctx, tVec.data(), NULL))
		die_ntc("finish");
	double eEst = 0.0;
	if (ntc_hll_estimate(tVec.data(), opt::nBits, 1, &eEst)) // nthll.cpp:243-252
		die_ntc("estimate");
	ntc_destroy(ctx);
	std::cout << "F0, Exp# of distnt kmers(k=" << opt::kmLen << "): " << (unsigned long long)eEst << "\n"; // nthll.cpp:254
	return 0;
}
