// The persistent tail kernel (tail_kernel.cuh) and its launcher, in a translation unit of its own so that it
// builds in parallel with the per-degree round kernels.
#define ZKSC_TAIL_IMPL
#include "tail_kernel.cuh"
using namespace zksc;

void zksc_launch_tail(dim3 grid, cudaStream_t s, const TailArgs& a) { tail_kernel<<<grid, kTailThreads, 0, s>>>(a); }

#ifdef ZKSC_TAIL_TRACE
// debug builds only (not declared in include/zksc.h): copy the timeline out, 64 rounds x 2 CTAs x 8 phases of %globaltimer ns
extern "C" int zksc_debug_tail_trace(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_tail_trace, sizeof(g_tail_trace));
}
#endif
