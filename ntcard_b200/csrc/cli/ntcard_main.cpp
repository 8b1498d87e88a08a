// ntcard -- host command line of the B200-native ntCard sketch path.
//
// Same surface as the reference's main() (ntcard.cpp:20-87, 317-478): `ntcard [OPTION]... FILE(S)...`,
// options -t -k -g -c -p -o (+ hidden -s -r, ignored -l -f), `@list` arguments, the "sBits = 7 below
// 50 GB" rule, format sniffing, <prefix>_k<k>.hist / compact output, "Runtime(sec)".  What differs is
// the inside: reader threads parse FASTQ/FASTA/SAM, split sequences at non-ACGTU characters, 2-bit
// pack them into pinned double buffers and hand batches to the device through the C-ABI
// (include/ntcard_b200.h); the hash -> sample -> increment loop and the counter-value histogram run
// on the GPU; the estimator runs on the host in the reference's arithmetic order.
#include <atomic>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <getopt.h>
#include <iomanip>
#include <iostream>
#include <future>
#include <mutex>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "ntcard_b200.h"
#include "reader.h"

#define PROGRAM "ntCard"

static const char VERSION_MESSAGE[] = PROGRAM " 1.2.2 (ntcard-b200: sm_100a sketch path)\n";

static const char USAGE_MESSAGE[] =
    "Usage: " PROGRAM " [OPTION]... FILE(S)...\n"
    "Estimates k-mer coverage histogram in FILE(S).\n\n"
    "Acceptable file formats: fastq, fasta, sam, and in compressed formats gz, bz2, xz.\n"
    "A list of files containing file names in each row can be passed with @ prefix.\n"
    "\n"
    " Options:\n"
    "\n"
    "  -t, --threads=N	use N parallel reader threads [1] (N>=2 should be used when input files are >=2)\n"
    "  -k, --kmer=N	the length of kmer (comma separated list for several k in one pass)\n"
    "  -g, --gap=N	the length of gap in the gap seed [0]. g mod 2 must equal k mod 2 unless g == 0\n"
    "           	-g does not support multiple k currently.\n"
    "  -c, --cov=N	the maximum coverage of kmer in output [1000]\n"
    "  -p, --pref=STRING    the prefix for output file name(s)\n"
    "  -o, --output=STRING	the name for output file name (used when output should be a single file)\n"
    "      --gpu=N	CUDA device to use [0]\n"
    "      --help	display this help and exit\n"
    "      --version	output version information and exit\n"
    "\n";

namespace opt {
unsigned nThrd = 1;
unsigned gap = 0;
unsigned rBits = 27;  // ntcard.cpp:57
unsigned sBits = 11;  // ntcard.cpp:58
unsigned covMax = 1000;
int gpu = 0;
int kernel = NTC_KERNEL_AUTO;
std::string prefix;
std::string output;
} // namespace opt

static const char shortopts[] = "t:s:r:k:c:l:p:f:o:g:";
enum { OPT_HELP = 1, OPT_VERSION, OPT_GPU, OPT_KERNEL };
static const struct option longopts[] = { { "threads", required_argument, NULL, 't' }, { "kmer", required_argument, NULL, 'k' },
	{ "gap", required_argument, NULL, 'g' }, { "cov", required_argument, NULL, 'c' }, { "rbit", required_argument, NULL, 'r' },
	{ "sbit", required_argument, NULL, 's' }, { "output", required_argument, NULL, 'o' }, { "pref", required_argument, NULL, 'p' },
	{ "gpu", required_argument, NULL, OPT_GPU }, { "kernel", required_argument, NULL, OPT_KERNEL },
	{ "help", no_argument, NULL, OPT_HELP }, { "version", no_argument, NULL, OPT_VERSION }, { NULL, 0, NULL, 0 } };

static void die_ntc(const char* what)
{
	std::cerr << PROGRAM ": " << what << ": " << ntc_last_error() << "\n";
	exit(EXIT_FAILURE);
}

// size on disk, as getInf (ntcard.cpp:89-94): compressed size for compressed inputs
static uint64_t file_size(const std::string& path)
{
	struct stat st;
	if (stat(path.c_str(), &st) != 0)
		return (uint64_t)-1; // the reference's tellg() on a failed stream is -1 as well
	return (uint64_t)st.st_size;
}

// outDefault, ntcard.cpp:277-298
static void write_default(const std::vector<unsigned>& kList, const uint64_t* totalKmers, const uint32_t* p_hist)
{
	for (unsigned k = 0; k < kList.size(); k++) {
		std::stringstream hstm;
		hstm << opt::prefix << "_k" << kList[k] << ".hist";
		std::ofstream hist(hstm.str().c_str());
		double F0 = 0.0;
		std::vector<double> f(opt::covMax + 1, 0.0);
		if (ntc_estimate(p_hist + (size_t)k * 2 * 65536, NULL, opt::rBits, opt::sBits, opt::covMax, &F0, f.data()))
			die_ntc("estimate");
		hist << "F1\t" << totalKmers[k] << "\n";
		hist << "F0\t" << (uint64_t)F0 << "\n";
		for (size_t i = 1; i <= opt::covMax; i++)
			hist << i << "\t" << (uint64_t)f[i] << "\n";
	}
}

// outCompact, ntcard.cpp:300-315
static void write_compact(const std::vector<unsigned>& kList, const uint64_t* totalKmers, const uint32_t* p_hist)
{
	std::ofstream hist(opt::output.c_str());
	hist << "k\tf\tn\n";
	for (unsigned k = 0; k < kList.size(); k++) {
		double F0 = 0.0;
		std::vector<double> f(opt::covMax + 1, 0.0);
		if (ntc_estimate(p_hist + (size_t)k * 2 * 65536, NULL, opt::rBits, opt::sBits, opt::covMax, &F0, f.data()))
			die_ntc("estimate");
		std::cerr << "k=" << kList[k] << "\tF1\t" << totalKmers[k] << "\n";
		std::cerr << "k=" << kList[k] << "\tF0\t" << (uint64_t)F0 << "\n";
		for (size_t i = 1; i <= opt::covMax; i++)
			hist << kList[k] << "\t" << i << "\t" << (uint64_t)f[i] << "\n";
	}
}

int main(int argc, char** argv)
{
	const auto t_start = std::chrono::steady_clock::now();
	std::vector<unsigned> kList;
	bool die = false;
	for (int c; (c = getopt_long(argc, argv, shortopts, longopts, NULL)) != -1;) {
		std::istringstream arg(optarg != NULL ? optarg : "");
		switch (c) {
		case '?': die = true; break;
		case 't': arg >> opt::nThrd; break;
		case 's': arg >> opt::sBits; break;
		case 'r': arg >> opt::rBits; break;
		case 'c':
			arg >> opt::covMax;
			if (opt::covMax > 65535)
				opt::covMax = 65535; // ntcard.cpp:342-343
			break;
		case 'p': arg >> opt::prefix; break;
		case 'o': arg >> opt::output; break;
		case 'g': arg >> opt::gap; break;
		case 'l':
		case 'f': { // accepted and ignored, like ntcard.cpp:69
			std::string ignored;
			arg >> ignored;
			break;
		}
		case 'k': {
			std::string token;
			while (getline(arg, token, ',')) {
				unsigned myK = 0;
				std::stringstream ss(token);
				ss >> myK;
				kList.push_back(myK);
			}
			break;
		}
		case OPT_GPU: arg >> opt::gpu; break;
		case OPT_KERNEL: arg >> opt::kernel; break;
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
	if (kList.empty()) {
		std::cerr << PROGRAM ": missing argument -k ... \n";
		die = true;
	}
	if (opt::prefix.empty() && opt::output.empty()) {
		std::cerr << PROGRAM ": missing argument -p/-o ... \n";
		die = true;
	}
	if (opt::gap != 0 && !kList.empty() && (opt::gap % 2 != kList[0] % 2)) { // ntcard.cpp:382-385
		std::cerr << PROGRAM "Gap size and kmer must have the same modulus\n";
		die = true;
	}
	if (opt::gap != 0 && kList.size() != 1) { // ntcard.cpp:397-400
		std::cerr << PROGRAM ": -g does not support multiple k currently.\n";
		die = true;
	}
	if (opt::gap != 0 && !kList.empty() && opt::gap + 2 > kList[0]) {
		std::cerr << PROGRAM ": -g must be smaller than k.\n";
		die = true;
	}
	if (kList.size() > NTC_MAX_K) {
		std::cerr << PROGRAM ": at most " << NTC_MAX_K << " k values per run.\n";
		die = true;
	}
	if (die) {
		std::cerr << "Try `" << PROGRAM << " --help' for more information.\n";
		exit(EXIT_FAILURE);
	}

	std::vector<std::string> inFiles; // ntcard.cpp:415-425
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
	uint64_t totalSize = 0; // ntcard.cpp:427-431
	for (auto& f : inFiles)
		totalSize += file_size(f);
	if (totalSize < 50000000000ULL)
		opt::sBits = 7;

	const bool timing = getenv("NTC_CLI_TIMING") != NULL; // phase times on stderr (not part of the reference's output)
	auto lap = [&](const char* what) {
		if (timing)
			std::cerr << "  [timing] " << what << ": " << std::setprecision(3) << std::fixed
			          << std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() << " s since start\n";
	};
	lap("options parsed");
	// The device context (CUDA initialisation + the sketch, ntcard.cpp:437-439) is created on a side thread WHILE the readers
	// already parse and pack: they only need it when their first batch is full.
	std::promise<ntc_ctx*> ctx_promise;
	std::shared_future<ntc_ctx*> ctx_future = ctx_promise.get_future().share();
	std::thread creator([&]() {
		ntc_ctx* c = NULL;
		if (ntc_create(&c, kList.data(), (unsigned)kList.size(), opt::rBits, opt::sBits, opt::gpu, NULL, NULL))
			die_ntc("cannot create the device sketch");
		if (opt::kernel != NTC_KERNEL_AUTO && ntc_set_kernel(c, opt::kernel))
			die_ntc("kernel");
		if (opt::gap != 0 && ntc_set_gap(c, opt::gap)) // stRead instead of ntRead, ntcard.cpp:184-185, 204-205, 231-232
			die_ntc("gap seed");
		lap("device context created (side thread)");
		ctx_promise.set_value(c);
	});
	unsigned kmin = kList[0];
	for (unsigned k : kList)
		kmin = k < kmin ? k : kmin;

	// reader threads: one file at a time per thread, dynamic (ntcard.cpp:445 `omp parallel for schedule(dynamic)`)
	std::atomic<size_t> next_file(0);
	std::mutex submit_mu;
	unsigned nthreads = opt::nThrd < 1 ? 1 : opt::nThrd;
	if (nthreads > inFiles.size())
		nthreads = (unsigned)inFiles.size();
	// -t larger than the number of files: the spare threads parse pieces of the same (uncompressed FASTQ / FASTA) file
	const unsigned per_file = (opt::nThrd < 1 ? 1u : opt::nThrd) / (nthreads ? nthreads : 1u);
	const std::function<ntc_ctx*()> ctx_source = [ctx_future]() { return ctx_future.get(); };
	std::atomic<int> worker_id(0);
	auto worker = [&]() {
		const auto w0 = std::chrono::steady_clock::now();
		ntcb::BatchSubmitter sub([ctx_future]() { return ctx_future.get(); }, kmin, &submit_mu);
		const auto w1 = std::chrono::steady_clock::now();
		struct Report {
			bool on;
			int id;
			std::chrono::steady_clock::time_point w0, w1;
			ntcb::BatchSubmitter* sub;
			~Report()
			{
				if (on) {
					const auto w2 = std::chrono::steady_clock::now();
					std::cerr << "  [timing] reader " << id << ": pinned buffers " << std::setprecision(3) << std::fixed
					          << std::chrono::duration<double>(w1 - w0).count() << " s, files " << std::chrono::duration<double>(w2 - w1).count()
					          << " s of which inside ntc_submit/ntc_wait " << sub->seconds_in_submit() << " s\n";
				}
			}
		} report{ timing, worker_id.fetch_add(1), w0, w1, &sub };
		for (;;) {
			size_t i = next_file.fetch_add(1);
			if (i >= inFiles.size())
				break;
			bool handled = false;
			if (per_file > 1)
				ntcb::read_file_parallel(inFiles[i], ctx_source, kmin, &submit_mu, per_file, &handled);
			if (!handled && !ntcb::read_file(inFiles[i], sub)) {
				std::cerr << "Error in reading file: " << inFiles[i] << std::endl; // ntcard.cpp:459-462
				exit(EXIT_FAILURE);
			}
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

	creator.join();
	ntc_ctx* ctx = ctx_future.get();
	lap("files read, packed and submitted");
	std::vector<uint64_t> totalKmers(kList.size(), 0);
	std::vector<uint32_t> p_hist(kList.size() * 2 * 65536);
	if (ntc_finish(ctx, NULL, totalKmers.data(), p_hist.data()))
		die_ntc("finish");
	lap("finish (flush + histogram)");
	if (opt::output.empty())
		write_default(kList, totalKmers.data(), p_hist.data());
	else
		write_compact(kList, totalKmers.data(), p_hist.data());
	lap("estimates written");
	ntc_destroy(ctx);
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
	std::cerr << "Runtime(sec): " << std::setprecision(4) << std::fixed << secs << "\n"; // ntcard.cpp:476
	return 0;
}
