// The resident rounds kernel (resident_kernel.cuh) and its launcher, split over three objects so that they build in parallel with
// the per-degree round kernels (compile with -DZKSC_RES_PART=0|1|2).  Instantiations: one per degree 1..5 for handles whose
// products all have that degree (registers allocated for that degree alone), plus the mixed-degree form (index 0).
//   part 0: degrees 1-3 and the dispatchers;  part 1: degrees 4, 5;  part 2: mixed degrees
#define ZKSC_RES_IMPL
#include "resident_kernel.cuh"
using namespace zksc;
#ifndef ZKSC_RES_PART
#error "compile with -DZKSC_RES_PART=<0|1|2>"
#endif

template <int DSEL>
static cudaError_t launch(unsigned int ctas, cudaStream_t s, const ResArgs& a) {
    void* params[] = {(void*)&a};
    // cooperative: the launch fails instead of starting a grid whose CTAs cannot all be resident (they wait for each other)
    return cudaLaunchCooperativeKernel((const void*)resident_kernel<DSEL>, dim3(ctas), dim3(kResThreads), params, 0, s);
}
template <int DSEL>
static int occ() {
    int o = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, resident_kernel<DSEL>, kResThreads, 0);
    return o;
}
cudaError_t zksc_launch_resident_hi(int dsel, unsigned int ctas, cudaStream_t s, const ResArgs& a);
cudaError_t zksc_launch_resident_mixed(unsigned int ctas, cudaStream_t s, const ResArgs& a);
int zksc_resident_occ_hi(int dsel);
int zksc_resident_occ_mixed();

#if ZKSC_RES_PART == 0
cudaError_t zksc_launch_resident(int dsel, unsigned int ctas, cudaStream_t s, const ResArgs& a) {
    switch (dsel) {
        case 1: return launch<1>(ctas, s, a);
        case 2: return launch<2>(ctas, s, a);
        case 3: return launch<3>(ctas, s, a);
        case 4: case 5: return zksc_launch_resident_hi(dsel, ctas, s, a);
        default: return zksc_launch_resident_mixed(ctas, s, a);
    }
}
// resident CTAs per SM of instantiation dsel
int zksc_resident_occ(int dsel) {
    switch (dsel) {
        case 1: return occ<1>();
        case 2: return occ<2>();
        case 3: return occ<3>();
        case 4: case 5: return zksc_resident_occ_hi(dsel);
        default: return zksc_resident_occ_mixed();
    }
}
#elif ZKSC_RES_PART == 1
cudaError_t zksc_launch_resident_hi(int dsel, unsigned int ctas, cudaStream_t s, const ResArgs& a) { return dsel == 4 ? launch<4>(ctas, s, a) : launch<5>(ctas, s, a); }
int zksc_resident_occ_hi(int dsel) { return dsel == 4 ? occ<4>() : occ<5>(); }
#else
cudaError_t zksc_launch_resident_mixed(unsigned int ctas, cudaStream_t s, const ResArgs& a) { return launch<0>(ctas, s, a); }
int zksc_resident_occ_mixed() { return occ<0>(); }
#endif

#if defined(ZKSC_RES_TRACE) && ZKSC_RES_PART == 0
// debug builds only (not declared in include/zksc.h): copy the timeline out, 64 rounds x 8 phases of %globaltimer ns (degree 2 instantiation)
extern "C" int zksc_debug_res_trace(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_res_trace, sizeof(g_res_trace)); }
extern "C" int zksc_debug_res_cta_trace(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_res_cta_trace, sizeof(g_res_cta_trace)); }
#endif
