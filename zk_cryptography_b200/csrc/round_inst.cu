// One translation unit per product degree (compile with -DZKSC_D=<1..8>): the two round kernels of that
// degree plus their host-side launcher and occupancy query.  Split out so the 16 heavy instantiations
// build in parallel.
#include "kernels.cuh"
#ifndef ZKSC_D
#error "compile with -DZKSC_D=<degree>"
#endif
#define ZKSC_CAT2(a, b) a##b
#define ZKSC_CAT(a, b) ZKSC_CAT2(a, b)
using namespace zksc;

void ZKSC_CAT(zksc_launch_round_, ZKSC_D)(bool fold, dim3 grid, cudaStream_t s, const RoundArgs& a) {
    if (fold) round_kernel<ZKSC_D, true><<<grid, kThreads, 0, s>>>(a);
    else round_kernel<ZKSC_D, false><<<grid, kThreads, 0, s>>>(a);
}
int ZKSC_CAT(zksc_occ_round_, ZKSC_D)(bool fold) {
    int o = 0;
    if (fold) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_kernel<ZKSC_D, true>, kThreads, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_kernel<ZKSC_D, false>, kThreads, 0);
    return o > 0 ? o : 1;
}
