// bitslice_inst.cu -- one translation unit per k mod 31 (compiled with -DBS_KM=<0..30>), so the unrolled variants
// of the scan kernel and of the fused sketch kernel build in parallel.  Each instantiates the kernel for sBits = 7 and 11, the two values the
// reference can reach without its hidden -s flag (ntcard.cpp:58, 430-431).
#include "fused_kernel.cuh"
#include "scan_kernel.cuh"

#ifndef BS_KM
#error "compile with -DBS_KM=<k mod 31>"
#endif

#define BS_CAT2(a, b) a##b
#define BS_CAT(a, b) BS_CAT2(a, b)

namespace ntc {
namespace pl {

cudaError_t BS_CAT(launch_scan_km_, BS_KM)(unsigned sBits, const ScanArgs& a)
{
	if (sBits == 7)
		return launch_scan_one<BS_KM, 7>(a);
	if (sBits == 11)
		return launch_scan_one<BS_KM, 11>(a);
	return cudaErrorInvalidValue;
}

// nthll pre-filter variants: candidates = k-mers whose canonical hash has its top T bits zero
cudaError_t BS_CAT(launch_hllscan_km_, BS_KM)(unsigned T, const ScanArgs& a)
{
	if (T == 0) // the filter follows the smallest register on the device (a.hll_min)
		return a.hll_min ? launch_scan_one<BS_KM, 13, 2>(a) : cudaErrorInvalidValue;
	if (T == 13)
		return launch_scan_one<BS_KM, 13, 1>(a);
	return cudaErrorInvalidValue;
}

cudaError_t BS_CAT(launch_fused_km_, BS_KM)(unsigned sBits, const FusedArgs& a)
{
	if (sBits == 7)
		return launch_fused_one<BS_KM, 7>(a);
	if (sBits == 11)
		return launch_fused_one<BS_KM, 11>(a);
	return cudaErrorInvalidValue;
}

} // namespace pl
} // namespace ntc
