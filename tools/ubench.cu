// Integer-pipe microbenchmarks for sm_100a: issue rates of the instructions the field arithmetic is
// made of, and whole-multiplication throughput.  SURVEY.md 8(d): "the integer roof must be measured by
// the builder".  Run on the GPU box:  build/ubench   (prints one JSON object per line)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../zk_cryptography_b200/csrc/fr.cuh"
using namespace zksc;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

enum { OP_IMAD = 0, OP_IMAD_HI = 1, OP_IMAD_WIDE = 2, OP_WIDE_CHAIN = 3, OP_IADD3 = 4, OP_IADD3_CARRY = 5, OP_LOHI_SPLIT = 6 };

template <int OP>
__global__ void __launch_bounds__(256) op_kernel(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 977u + seed;
    uint32_t c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = a + i;
    for (int it = 0; it < iters; it++) {
        if (OP == OP_IMAD) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a), "r"(b));
        } else if (OP == OP_IMAD_HI) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a), "r"(b));
        } else if (OP == OP_IMAD_WIDE) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long w = ((unsigned long long)c[i + 1] << 32) | c[i];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(a), "r"(b));
                c[i] = (uint32_t)w; c[i + 1] = (uint32_t)(w >> 32);
            }
        } else if (OP == OP_WIDE_CHAIN) {
            // two carry chains of 4 fused lo/hi pairs each (what chain4 emits): 8 IMAD.WIDE(.X) + 2 IADD3.X
#pragma unroll
            for (int h = 0; h < 16; h += 8) {
                asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                             "madc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                             "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                             "madc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.cc.u32 %7, %8, %9, %7;\n\t"
                             "addc.u32 %8, %8, 0;"
                             : "+r"(c[h]), "+r"(c[h + 1]), "+r"(c[h + 2]), "+r"(c[h + 3]), "+r"(c[h + 4]), "+r"(c[h + 5]), "+r"(c[h + 6]), "+r"(c[h + 7]), "+r"(a)
                             : "r"(b));
            }
        } else if (OP == OP_IADD3) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(b));
        } else if (OP == OP_IADD3_CARRY) {
#pragma unroll
            for (int h = 0; h < 16; h += 8)
                asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %8;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %8;\n\t"
                             "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %8;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %8;"
                             : "+r"(c[h]), "+r"(c[h + 1]), "+r"(c[h + 2]), "+r"(c[h + 3]), "+r"(c[h + 4]), "+r"(c[h + 5]), "+r"(c[h + 6]), "+r"(c[h + 7])
                             : "r"(b));
        } else if (OP == OP_LOHI_SPLIT) {
            // the unfused form ptxas picks when the multiplier depends on the accumulator: IMAD + IMAD.HI
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                asm volatile("mad.lo.u32 %0, %2, %3, %0;\n\tmad.hi.u32 %1, %2, %3, %1;" : "+r"(c[i]), "+r"(c[i + 1]) : "r"(a), "r"(b));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s ^ a;
}

// MODE 0: x = fr_mul(x, y)  (dependent chain per thread; TLP hides latency)
// MODE 1: acc += mul_wide(x, y) ; x perturbed (lazy product + 17-limb accumulate)
// MODE 2: x = fr_fold(x, y, r)
// MODE 3: x = fr_add(x, y)
template <int MODE>
__global__ void __launch_bounds__(128) fr_kernel(const Fr* in, Fr* out, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fr x = ld256(in + (tid & 1023)), y = ld256(in + ((tid + 1) & 1023)), r = ld256(in + ((tid + 2) & 1023));
    Acc<17> acc; acc_zero(acc);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) x = fr_mul(x, y);
        else if (MODE == 1) { uint32_t T[16]; mul_wide(T, x, y); acc_add<17, 16>(acc, T); x.l[0] ^= T[3]; }
        else if (MODE == 2) x = fr_fold(x, y, r);
        else x = fr_add(x, y);
    }
    if (MODE == 1) x = acc17_reduce(acc);
    st256(out + tid, x);
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
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d}\n", prop.name, sms, clk_khz / 1000);
    uint32_t* out; CHECK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 4));
    const int iters = 4096, blocks = sms * 8;
    const char* names[] = {"imad_lo", "imad_hi", "imad_wide", "wide_carry_chain(8 wide + 2 iadd3.x)", "iadd3", "iadd3_carry_chain", "imad_lo+imad_hi pair"};
    double ops_per_iter[] = {16, 16, 8, 8, 16, 16, 16};
    for (int op = 0; op < 7; op++) {
        double ms = 0;
        auto L = [&]() {
            switch (op) {
                case 0: op_kernel<0><<<blocks, 256>>>(out, iters, 1); break;
                case 1: op_kernel<1><<<blocks, 256>>>(out, iters, 1); break;
                case 2: op_kernel<2><<<blocks, 256>>>(out, iters, 1); break;
                case 3: op_kernel<3><<<blocks, 256>>>(out, iters, 1); break;
                case 4: op_kernel<4><<<blocks, 256>>>(out, iters, 1); break;
                case 5: op_kernel<5><<<blocks, 256>>>(out, iters, 1); break;
                default: op_kernel<6><<<blocks, 256>>>(out, iters, 1); break;
            }
        };
        ms = time_ms(L, 5);
        double total = (double)blocks * 256 * iters * ops_per_iter[op];
        printf("{\"bench\": \"%s\", \"ms\": %.4f, \"Gops_per_s\": %.1f, \"ops_per_clk_per_sm_at_max_clock\": %.2f}\n", names[op], ms, total / ms / 1e6,
               total / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    Fr *fin, *fout; CHECK(cudaMalloc(&fin, 1024 * sizeof(Fr))); CHECK(cudaMalloc(&fout, (size_t)sms * 16 * 128 * sizeof(Fr)));
    { // deterministic inputs < r: limbs with top limb small
        Fr h[1024];
        for (int i = 0; i < 1024; i++) for (int k = 0; k < 8; k++) h[i].l[k] = (k == 7) ? (0x1234567u + i) : (2654435761u * (i * 8 + k + 1));
        CHECK(cudaMemcpy(fin, h, sizeof h, cudaMemcpyHostToDevice));
    }
    const char* fnames[] = {"fr_mul (full Montgomery product)", "mul_wide + 17-limb accumulate (lazy product)", "fr_fold (sub + mul + add)", "fr_add"};
    for (int wpb = 0; wpb < 2; wpb++) {
        const int fblocks = sms * (wpb ? 16 : 4), fiters = 2048;
        for (int mode = 0; mode < 4; mode++) {
            auto L = [&]() {
                switch (mode) {
                    case 0: fr_kernel<0><<<fblocks, 128>>>(fin, fout, fiters); break;
                    case 1: fr_kernel<1><<<fblocks, 128>>>(fin, fout, fiters); break;
                    case 2: fr_kernel<2><<<fblocks, 128>>>(fin, fout, fiters); break;
                    default: fr_kernel<3><<<fblocks, 128>>>(fin, fout, fiters); break;
                }
            };
            double ms = time_ms(L, 5);
            double total = (double)fblocks * 128 * fiters;
            printf("{\"bench\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gops_per_s\": %.2f, \"clk_per_op_per_sm_at_max_clock\": %.2f}\n", fnames[mode], fblocks / sms, ms,
                   total / ms / 1e6, (ms * 1e-3) * (clk_khz * 1e3) * sms / total);
        }
    }
    CHECK(cudaGetLastError());
    return 0;
}
