// The persistent tail kernel (tail_kernel.cuh) and its launcher, in a translation unit of its own so that it
// builds in parallel with the per-degree round kernels.
#define ZKSC_TAIL_IMPL
#include "tail_kernel.cuh"
using namespace zksc;

void zksc_launch_tail(dim3 grid, cudaStream_t s, const TailArgs& a) { tail_kernel<<<grid, kTailThreads, 0, s>>>(a); }
