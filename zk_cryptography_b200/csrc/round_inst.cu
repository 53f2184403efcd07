// One translation unit per product degree (compile with -DZKSC_D=<1..8>): the three round kernels of that
// degree plus their host-side launcher and occupancy query.  Split out so the 24 heavy instantiations
// build in parallel.
#include "kernels.cuh"
#ifndef ZKSC_D
#error "compile with -DZKSC_D=<degree>"
#endif
#define ZKSC_CAT2(a, b) a##b
#define ZKSC_CAT(a, b) ZKSC_CAT2(a, b)
using namespace zksc;

// variant: 0 = evaluate only, 1 = fold + evaluate, 2 = fold + evaluate without point 1 (derived on the host)
void ZKSC_CAT(zksc_launch_round_, ZKSC_D)(int variant, dim3 grid, cudaStream_t s, const RoundArgs& a) {
    if (variant == 2) round_kernel<ZKSC_D, true, true><<<grid, kThreads, 0, s>>>(a);
    else if (variant == 1) round_kernel<ZKSC_D, true, false><<<grid, kThreads, 0, s>>>(a);
    else round_kernel<ZKSC_D, false, false><<<grid, kThreads, 0, s>>>(a);
}
int ZKSC_CAT(zksc_occ_round_, ZKSC_D)(int variant) {
    int o = 0;
    if (variant == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_kernel<ZKSC_D, true, true>, kThreads, 0);
    else if (variant == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_kernel<ZKSC_D, true, false>, kThreads, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_kernel<ZKSC_D, false, false>, kThreads, 0);
    return o > 0 ? o : 1;
}
