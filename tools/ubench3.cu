// Third-round microbenchmarks: whole-multiplication variants and the FP64 / IMAD.HI issue rates.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../zk_cryptography_b200/csrc/fr.cuh"
#include "fr_rowwise.cuh"
using namespace zksc;
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// MODE 0: first-generation row-wise multiplier   MODE 1: library fr_mul   MODE 4: product-scanning multiplier
// MODE 2: mul_wide + acc17                     MODE 3: mul_ps<false> + acc17
template <int MODE>
__global__ void __launch_bounds__(128) fr_kernel(const Fr* in, Fr* out, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fr x = ld256(in + (tid & 1023)), y = ld256(in + ((tid + 1) & 1023));
    Acc<17> acc; acc_zero(acc);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) x = rowwise::fr_mul(x, y);
        else if (MODE == 1) x = fr_mul(x, y);
        else if (MODE == 2) { uint32_t T[16]; rowwise::mul_wide(T, x, y); acc_add<17, 16>(acc, T); x.l[0] ^= T[3]; }
        else if (MODE == 3) { uint32_t T[16]; (void)mul_ps<false>(T, x, y); acc_add<17, 16>(acc, T); x.l[0] ^= T[3]; }
        else { Fr o; (void)mul_ps<true>(o.l, x, y); cond_sub_r(o.l); x = o; }
    }
    if (MODE == 2 || MODE == 3) x = acc17_reduce(acc);
    st256(out + tid, x);
}

// OP 0: IMAD.HI.U32 (8 indep, varying operands)   OP 1: IMAD lo   OP 2: DFMA.RZ (8 indep)   OP 3: DFMA + IMAD.WIDE interleaved
template <int OP>
__global__ void __launch_bounds__(256) op_kernel(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed;
    uint32_t c[16];
    double d[8];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = a + i * 77u;
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = 1.0 + (double)(a & 1023) * (i + 1) * 1e-9;
    for (int it = 0; it < iters; it++) {
        if (OP == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(c[(i + 3) & 15]), "r"(c[(i + 7) & 15]));
        } else if (OP == 1) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(c[(i + 3) & 15]), "r"(c[(i + 7) & 15]));
        } else if (OP == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[(i + 3) & 7]), "d"(d[(i + 5) & 7]));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[(i + 3) & 7]), "d"(d[(i + 5) & 7]));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(c[2 * i]), "+r"(c[2 * i + 1]) : "r"(c[(2 * i + 3) & 15]), "r"(c[(2 * i + 6) & 15]));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= c[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= (uint32_t)__double_as_longlong(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    uint32_t* out; CHECK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 4));
    const int iters = 4096, blocks = sms * 8;
    const char* names[] = {"IMAD.HI.U32 (16 indep)", "IMAD lo (16 indep)", "DFMA.RZ (8 indep)", "8 DFMA.RZ + 8 IMAD.WIDE interleaved"};
    const double per_iter[] = {16, 16, 8, 8};
    for (int op = 0; op < 4; op++) {
        auto L = [&]() {
            switch (op) {
                case 0: op_kernel<0><<<blocks, 256>>>(out, iters, 1); break;
                case 1: op_kernel<1><<<blocks, 256>>>(out, iters, 1); break;
                case 2: op_kernel<2><<<blocks, 256>>>(out, iters, 1); break;
                default: op_kernel<3><<<blocks, 256>>>(out, iters, 1); break;
            }
        };
        double ms = time_ms(L, 5);
        double total = (double)blocks * 256 * iters * per_iter[op];
        printf("{\"bench\": \"%s\", \"ms\": %.4f, \"ops_per_clk_per_sm_at_max_clock\": %.2f}\n", names[op], ms, total / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    Fr *fin, *fout; CHECK(cudaMalloc(&fin, 1024 * sizeof(Fr))); CHECK(cudaMalloc(&fout, (size_t)sms * 16 * 128 * sizeof(Fr)));
    {
        static Fr h[1024];
        for (int i = 0; i < 1024; i++) for (int k = 0; k < 8; k++) h[i].l[k] = (k == 7) ? (0x1234567u + i) : (2654435761u * (i * 8 + k + 1));
        CHECK(cudaMemcpy(fin, h, sizeof h, cudaMemcpyHostToDevice));
    }
    const char* fnames[] = {"fr_mul row-wise, negated digit (first generation)", "fr_mul row-wise, complement digit (library)", "mul_wide + acc17 (lazy, row-wise, library)", "mul_ps<false> + acc17 (lazy, product scanning)", "mul_ps<true> product scanning, complement digit"};
    for (int wpb = 0; wpb < 2; wpb++) {
        const int fblocks = sms * (wpb ? 16 : 4), fiters = 2048;
        for (int mode = 0; mode < 5; mode++) {
            auto L = [&]() {
                switch (mode) {
                    case 0: fr_kernel<0><<<fblocks, 128>>>(fin, fout, fiters); break;
                    case 1: fr_kernel<1><<<fblocks, 128>>>(fin, fout, fiters); break;
                    case 2: fr_kernel<2><<<fblocks, 128>>>(fin, fout, fiters); break;
                    case 3: fr_kernel<3><<<fblocks, 128>>>(fin, fout, fiters); break;
                    default: fr_kernel<4><<<fblocks, 128>>>(fin, fout, fiters); break;
                }
            };
            double ms = time_ms(L, 5);
            double total = (double)fblocks * 128 * fiters;
            printf("{\"bench\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gops_per_s\": %.2f, \"clk_per_op_per_sm_at_max_clock\": %.2f}\n", fnames[mode],
                   fblocks / sms, ms, total / ms / 1e6, (ms * 1e-3) * (clk_khz * 1e3) * sms / total);
        }
    }
    CHECK(cudaGetLastError());
    return 0;
}
