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
// staged: the TMA-staged kernel (degrees <= 5, pairs per table a multiple of 32)
template <int NB>
static void launch(int variant, bool staged, dim3 grid, cudaStream_t s, const RoundArgsT<NB>& a) {
#if ZKSC_D <= 5
    if (staged) {
        if (variant == 2) round_tma_kernel<ZKSC_D, true, true, NB><<<grid, kThreads, kWarps * tma_slot_bytes<ZKSC_D, true>(), s>>>(a);
        else if (variant == 1) round_tma_kernel<ZKSC_D, true, false, NB><<<grid, kThreads, kWarps * tma_slot_bytes<ZKSC_D, true>(), s>>>(a);
        else round_tma_kernel<ZKSC_D, false, false, NB><<<grid, kThreads, kWarps * tma_slot_bytes<ZKSC_D, false>(), s>>>(a);
        return;
    }
#endif
    if (variant == 2) round_kernel<ZKSC_D, true, true, NB><<<grid, kThreads, 0, s>>>(a);
    else if (variant == 1) round_kernel<ZKSC_D, true, false, NB><<<grid, kThreads, 0, s>>>(a);
    else round_kernel<ZKSC_D, false, false, NB><<<grid, kThreads, 0, s>>>(a);
}
void ZKSC_CAT(zksc_launch_round_, ZKSC_D)(int variant, bool staged, dim3 grid, cudaStream_t s, const RoundArgsT<1>& a) { launch<1>(variant, staged, grid, s, a); }
void ZKSC_CAT(zksc_launch_round_, ZKSC_D)(int variant, bool staged, dim3 grid, cudaStream_t s, const RoundArgsT<kMaxBatch>& a) { launch<kMaxBatch>(variant, staged, grid, s, a); }
#if ZKSC_D <= 5
template <int NB>
static void allow_smem() {
    cudaFuncSetAttribute(round_tma_kernel<ZKSC_D, true, true, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * tma_slot_bytes<ZKSC_D, true>());
    cudaFuncSetAttribute(round_tma_kernel<ZKSC_D, true, false, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * tma_slot_bytes<ZKSC_D, true>());
    cudaFuncSetAttribute(round_tma_kernel<ZKSC_D, false, false, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * tma_slot_bytes<ZKSC_D, false>());
}
#endif
// Called once per context (device): opt the staged kernels into their dynamic shared memory.
void ZKSC_CAT(zksc_prepare_round_, ZKSC_D)() {
#if ZKSC_D <= 5
    allow_smem<1>();
    allow_smem<kMaxBatch>();
#endif
}
// resident CTAs per SM; variants 3..5 = the staged kernels of variants 0..2 (0 when there is none)
int ZKSC_CAT(zksc_occ_round_, ZKSC_D)(int variant) {
    int o = 0;
    if (variant >= 6) return 0;
    if (variant == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_kernel<ZKSC_D, true, true, 1>, kThreads, 0);
    else if (variant == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_kernel<ZKSC_D, true, false, 1>, kThreads, 0);
    else if (variant == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_kernel<ZKSC_D, false, false, 1>, kThreads, 0);
#if ZKSC_D <= 5
    else if (variant == 5) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_tma_kernel<ZKSC_D, true, true, 1>, kThreads, kWarps * tma_slot_bytes<ZKSC_D, true>());
    else if (variant == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_tma_kernel<ZKSC_D, true, false, 1>, kThreads, kWarps * tma_slot_bytes<ZKSC_D, true>());
    else if (variant == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, round_tma_kernel<ZKSC_D, false, false, 1>, kThreads, kWarps * tma_slot_bytes<ZKSC_D, false>());
#else
    else return 0;
#endif
    return o > 0 ? o : 1;
}
