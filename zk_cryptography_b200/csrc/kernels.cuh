// Sumcheck round kernels for sm_100a.
//
// One launch = one round of one product  prod_k f_k  of D multilinear tables, for `gridDim.y`
// independent proofs.  It replaces, per round, the reference's
//     for i in 0..=d: p.partial_evaluation(F::from(i), 0).element_wise_product().iter().sum()
// (sumcheck/src/composed/multi_composed_sumcheck.rs:81-89, composed/composed_sumcheck.rs:41-49,
//  and for D = 1 Multilinear::split_poly_into_two_and_sum_each_part, evaluation_form.rs:68-74)
// and, when FOLD is set, also the previous round's
//     current_poly[i].partial_evaluation(&random_r, &0)          (multi_composed_sumcheck.rs:103-105)
// so that every round is ONE pass over HBM: read T_{j-1} (4 entries per table per thread), write
// T_j (2 entries), and accumulate the d+1 evaluations of round j from the two freshly folded entries.
//
// SKIP1: evaluation point 1 is not computed: the host derives it as h(1) = claim - h(0), where claim is
// the previous round polynomial at the previous challenge (the identity the verifier checks), which is
// exact field arithmetic and therefore bit-identical -- one product fewer per pair.
//
// Variable 0 is the most significant index bit (polynomial/src/utils.rs:26-53 with index 0): round j
// pairs entry x with x + N_j/2.
#pragma once
#include "fr.cuh"

namespace zksc {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBatch = 64;   // proofs per launch (challenges travel as kernel parameters)
constexpr int kMaxDegree = 8;

struct RoundArgs {
    const Fr* in;              // table 0 of proof 0 (current tables)
    Fr* out;                   // FOLD: where the folded tables go (may alias `in`)
    unsigned long long in_tab_stride, in_proof_stride;    // elements
    unsigned long long out_tab_stride, out_proof_stride;  // elements
    unsigned long long half;   // pairs per table in the round being evaluated (N_j / 2)
    Fr* partials;              // [proof][block][npts] scratch
    unsigned int* counters;    // [proof], zero on entry, zero on exit
    Fr* result;                // [proof][res_stride]: npts Montgomery elements each
    unsigned int res_stride;
    unsigned int npts;         // evaluation points 0..npts-1 are wanted (<= D+1); SKIP1 kernels leave point 1 untouched
    volatile unsigned int* flag;  // optional: set to flag_value (system scope) after the results
    unsigned int flag_value;
    Fr chal[kMaxBatch];        // FOLD: challenge of the previous round, Montgomery form, per proof
};

template <int D>
struct Lazy {
    static constexpr bool wide = (D <= 3);
    static constexpr int NL = (D == 1) ? 9 : (wide ? 17 : 9);
};

// Degrees above 5 are rare (the reference's benches stop at 5): their multiplications are real calls, which
// keeps the code size -- and the build time -- of those instantiations in check.
static __device__ __noinline__ Fr fr_mul_call(const Fr& a, const Fr& b) { return fr_mul(a, b); }
static __device__ __noinline__ Fr fr_fold_call(const Fr& a, const Fr& b, const Fr& r) { return fr_fold(a, b, r); }
template <int D>
ZKSC_DEV Fr fr_mul_d(const Fr& a, const Fr& b) {
    if constexpr (D > 5) return fr_mul_call(a, b);
    else return fr_mul(a, b);
}
template <int D>
ZKSC_DEV Fr fr_fold_d(const Fr& a, const Fr& b, const Fr& r) {
    if constexpr (D > 5) return fr_fold_call(a, b, r);
    else return fr_fold(a, b, r);
}

// acc += prod_k f[k]
template <int D>
ZKSC_DEV void accumulate_product(Acc<Lazy<D>::NL>& acc, const Fr (&f)[D]) {
    if constexpr (D == 1) {
        acc_add<9, 8>(acc, f[0].l);
    } else if constexpr (Lazy<D>::wide) {
        Fr g = f[0];
#pragma unroll
        for (int k = 1; k < D - 1; k++) g = fr_mul(g, f[k]);
        uint32_t T[16];
        mul_wide(T, g, f[D - 1]);      // last multiplication stays unreduced
        acc_add<17, 16>(acc, T);
    } else {
        Fr g = f[0];
#pragma unroll
        for (int k = 1; k < D; k++) g = fr_mul_d<D>(g, f[k]);
        acc_add<9, 8>(acc, g.l);
    }
}

template <int NL>
ZKSC_DEV Fr acc_finish(const Acc<NL>& a) {
    if constexpr (NL == 9) return acc9_reduce(a);
    else return acc17_reduce(a);
}

// Block-level reduction of NP accumulators, publication of the block partial, and -- in the last
// block of each proof to arrive -- the final cross-block sum.
// Accumulator slot s holds evaluation point s, or with SKIP1 point (s == 0 ? 0 : s + 1).
template <int NL, int NP, bool SKIP1>
ZKSC_DEV void reduce_and_publish(Acc<NL> (&acc)[NP], const RoundArgs& args, int npts) {
    auto point_of = [](int slot) { return (SKIP1 && slot > 0) ? slot + 1 : slot; };
    __shared__ Acc<NL> s_warp[kWarps][NP];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int proof = blockIdx.y;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        acc_warp_reduce(acc[p]);
        if (lane == 0) s_warp[warp][p] = acc[p];
    }
    __syncthreads();
    Fr* my_partials = args.partials + ((size_t)proof * gridDim.x + blockIdx.x) * NP;
    if (warp == 0) {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            Acc<NL> a;
            if (lane < kWarps) a = s_warp[lane][p];
            else acc_zero(a);
            acc_warp_reduce(a);
            if (lane == 0 && point_of(p) < npts) {
                Fr v = acc_finish<NL>(a);
                st256(my_partials + p, v);
            }
        }
        if (lane == 0) {
            __threadfence();
            unsigned int done = atomicAdd(args.counters + proof, 1u);
            s_last = (done == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block of this proof: sum the per-block partials (canonical Montgomery elements)
    const Fr* all = args.partials + (size_t)proof * gridDim.x * NP;
    for (int p = warp; p < NP; p += kWarps) {
        if (point_of(p) >= npts) continue;
        Acc<9> a;
        acc_zero(a);
        for (unsigned int blk = lane; blk < gridDim.x; blk += 32) {
            Fr v = ld256(all + (size_t)blk * NP + p);
            acc_add<9, 8>(a, v.l);
        }
        acc_warp_reduce(a);
        if (lane == 0) {
            Fr v = acc9_reduce(a);
            st256(args.result + (size_t)proof * args.res_stride + point_of(p), v);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        args.counters[proof] = 0;
        if (args.flag) {
            __threadfence_system();
            // one flag word per proof would be wasteful: proofs bump a shared word via the counter slot
            unsigned int fin = atomicAdd(args.counters + gridDim.y, 1u);
            if (fin == gridDim.y - 1) {
                args.counters[gridDim.y] = 0;
                __threadfence_system();
                *args.flag = args.flag_value;
            }
        }
    }
}

// FOLD = false : evaluate the round polynomial of the tables as they are (first round).
// FOLD = true  : bind the previous challenge (in -> out), then evaluate the next round on the result.
template <int D, bool FOLD, bool SKIP1>
__global__ void __launch_bounds__(kThreads) round_kernel(const __grid_constant__ RoundArgs args) {
    constexpr int NL = Lazy<D>::NL;
    constexpr int NP = SKIP1 ? D : D + 1;   // accumulators
    const int proof = blockIdx.y;
    const Fr* in = args.in + (size_t)proof * args.in_proof_stride;
    Fr* out = args.out + (size_t)proof * args.out_proof_stride;
    const unsigned long long half = args.half;
    const int npts = args.npts;

    Fr r;
    if constexpr (FOLD) r = args.chal[proof];

    Acc<NL> acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) acc_zero(acc[p]);

    const unsigned long long stride = (unsigned long long)gridDim.x * kThreads;
    for (unsigned long long x = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; x < half; x += stride) {
        Fr a[D], b[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            const Fr* t = in + (size_t)k * args.in_tab_stride;
            if constexpr (FOLD) {
                // T_{j-1} has 4*half entries; its pairs are (y, y + 2*half)
                Fr p0 = ld256(t + x), p1 = ld256(t + x + 2 * half);
                Fr q0 = ld256(t + x + half), q1 = ld256(t + x + 3 * half);
                a[k] = fr_fold_d<D>(p0, p1, r);
                b[k] = fr_fold_d<D>(q0, q1, r);
                Fr* o = out + (size_t)k * args.out_tab_stride;
                st256(o + x, a[k]);
                st256(o + x + half, b[k]);
            } else {
                a[k] = ld256_stream(t + x);
                b[k] = ld256_stream(t + x + half);
            }
        }
        // evaluation points 0 and 1 are the two halves themselves
        accumulate_product<D>(acc[0], a);
        if (!SKIP1 && npts > 1) accumulate_product<D>(acc[1], b);
        if (D >= 2 && npts > 2) {
            // f_k(t) = a_k + t (b_k - a_k): walk t = 2..D by repeated addition of the difference
            Fr delta[D];
#pragma unroll
            for (int k = 0; k < D; k++) delta[k] = fr_sub(b[k], a[k]);
#pragma unroll
            for (int p = 2; p <= D; p++) {
#pragma unroll
                for (int k = 0; k < D; k++) b[k] = fr_add(b[k], delta[k]);
                if (p < npts) accumulate_product<D>(acc[SKIP1 ? p - 1 : p], b);
            }
        }
    }
    reduce_and_publish<NL, NP, SKIP1>(acc, args, npts);
}

}  // namespace zksc
