// Second-round integer microbenchmarks: which carry forms of IMAD.WIDE run at full rate on sm_100a?
//   build/ubench2   (one JSON object per line)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// OP 0: 8 independent IMAD.WIDE with carry-OUT only, carry counted by IADD3.X  (product scanning)
// OP 1: one chain of 8 lo/hi pairs (1 plain + 7 .X) + IADD3.X
// OP 2: 8 independent plain mad.wide.u32 (no carries)                               (reduced radix)
// OP 3: 8 independent IMAD.WIDE carry-out only, carries NOT consumed
// OP 4: OP 0 + 8 extra independent IADD3 (ALU pipe co-issue test)
// OP 5: 8 plain mad.wide + 8 independent IADD3 + 8 LOP3 (ALU co-issue with plain wide)
// OP 6: 8 plain mad.wide + 16 ALU ops
template <int OP>
__global__ void __launch_bounds__(256) op_kernel(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 977u + seed;
    uint32_t c[16], cnt[8], x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i] = a + i; x[i] = a ^ i; }
#pragma unroll
    for (int i = 0; i < 8; i++) cnt[i] = 0;
    for (int it = 0; it < iters; it++) {
        if (OP == 0 || OP == 4) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                             : "+r"(c[2 * i]), "+r"(c[2 * i + 1]), "+r"(cnt[i]) : "r"(c[(2 * i + 3) & 15]), "r"(c[(2 * i + 6) & 15]));
            if (OP == 4) {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
            }
        } else if (OP == 1) {
            asm volatile("mad.lo.cc.u32 %0, %17, %18, %0;\n\tmadc.hi.cc.u32 %1, %17, %18, %1;\n\t"
                         "madc.lo.cc.u32 %2, %17, %18, %2;\n\tmadc.hi.cc.u32 %3, %17, %18, %3;\n\t"
                         "madc.lo.cc.u32 %4, %17, %18, %4;\n\tmadc.hi.cc.u32 %5, %17, %18, %5;\n\t"
                         "madc.lo.cc.u32 %6, %17, %18, %6;\n\tmadc.hi.cc.u32 %7, %17, %18, %7;\n\t"
                         "madc.lo.cc.u32 %8, %17, %18, %8;\n\tmadc.hi.cc.u32 %9, %17, %18, %9;\n\t"
                         "madc.lo.cc.u32 %10, %17, %18, %10;\n\tmadc.hi.cc.u32 %11, %17, %18, %11;\n\t"
                         "madc.lo.cc.u32 %12, %17, %18, %12;\n\tmadc.hi.cc.u32 %13, %17, %18, %13;\n\t"
                         "madc.lo.cc.u32 %14, %17, %18, %14;\n\tmadc.hi.cc.u32 %15, %17, %18, %15;\n\t"
                         "addc.u32 %16, %16, 0;"
                         : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), "+r"(c[8]), "+r"(c[9]),
                           "+r"(c[10]), "+r"(c[11]), "+r"(c[12]), "+r"(c[13]), "+r"(c[14]), "+r"(c[15]), "+r"(cnt[0])
                         : "r"(x[0]), "r"(x[1]));
            x[0] += c[3]; x[1] ^= c[8];
        } else if (OP == 2 || OP == 5 || OP == 6) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(c[2 * i]), "+r"(c[2 * i + 1]) : "r"(c[(2 * i + 3) & 15]), "r"(c[(2 * i + 6) & 15]));
            if (OP == 5 || OP == 6) {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
#pragma unroll
                for (int i = 8; i < 16; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
            }
            if (OP == 6) {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(cnt[i]) : "r"(x[i]));
            }
        } else if (OP == 3) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(c[2 * i]), "+r"(c[2 * i + 1]) : "r"(c[(2 * i + 3) & 15]), "r"(c[(2 * i + 6) & 15]));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= c[i] ^ x[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s ^ a;
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
    const char* names[] = {"wide carry-out + IADD3.X count (8 indep)", "chain of 8 pairs (1 plain + 7 .X) + IADD3.X", "plain mad.wide (8 indep)",
                           "wide carry-out, carry unused (8 indep)", "wide carry-out + count + 8 IADD3", "plain wide + 8 IADD3 + 8 LOP3", "plain wide + 8 IADD3 + 8 LOP3 + 8 SHF"};
    for (int op = 0; op < 7; op++) {
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
        double ms = time_ms(L, 5);
        double total = (double)blocks * 256 * iters * 8;
        printf("{\"bench\": \"%s\", \"ms\": %.4f, \"wide_products_per_clk_per_sm_at_max_clock\": %.2f, \"clk_per_8_products_per_warp_per_smsp\": %.2f}\n", names[op], ms,
               total / (ms * 1e-3) / sms / (clk_khz * 1e3), (ms * 1e-3) * (clk_khz * 1e3) / ((double)blocks / sms * 8 / 4) / iters);
    }
    CHECK(cudaGetLastError());
    return 0;
}
