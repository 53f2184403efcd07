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
// RoundBase::counters: [group] arrivals | [n_groups] groups done | from kWorkCtrBase on, for launches of at most kDynMaxGroups groups: the
// work counters, one per (group, warp index in the CTA), each in a cache line of its own -- atomics on ONE address go through at
// ~0.8 per ns (measured: a degree-1 round of 2^26 pairs took 2.6 ms instead of 1.0 with a single counter per group)
constexpr unsigned int kDynMaxGroups = 64;
constexpr unsigned int kWorkCtrBase = kMaxBatch * 8 + 32, kWorkCtrWords = 32;
constexpr unsigned int kRoundCounterWords = kWorkCtrBase + kDynMaxGroups * 4 * kWorkCtrWords;
constexpr int kMaxDegree = 8;

constexpr int kMaxRanks = 8;    // GPUs of one NVSwitch box reachable through peer memory
// Cross-GPU exchange of a round's partial evaluations, fused into the tail of the round kernel (sharded
// contexts).  Every rank owns an exchange buffer  data[2][n_ranks][cap]  +  tags[2][n_ranks]  in its own HBM,
// opened by the peers through CUDA IPC.  The last block of the round stores this rank's partials into slot
// [seq & 1][rank] of EVERY rank's buffer (peer stores over NVLink), fences, stores the round's sequence
// number into the matching tag of every rank, waits until all n_ranks tags of its own buffer carry that
// number, adds the n_ranks partials mod r and writes the sums to the host-mapped result buffer.  No NCCL call,
// no extra launch, no host reduction.  Two slots suffice: a rank can be at most one round ahead of a peer,
// because it needs that peer's partials of round s before it can leave round s.
struct XchArgs {
    Fr* peer_data[kMaxRanks];
    unsigned int* peer_tags[kMaxRanks];
    const Fr* send;            // this rank's partials (n_elems elements, device memory)
    Fr* out;                   // where the reduced values go (host-mapped)
    unsigned int n_ranks, rank, seq, n_elems, cap;
};

struct RoundBase {
    const Fr* in;              // table 0 of proof 0 (current tables)
    Fr* out;                   // FOLD: where the folded tables go (may alias `in`)
    unsigned long long in_tab_stride, in_proof_stride;    // elements
    unsigned long long out_tab_stride, out_proof_stride;  // elements
    // gridDim.z > 1: the launch covers that many products of the SAME degree (blockIdx.z = product); element strides
    // between consecutive products' first tables, and between their result slots
    unsigned long long in_prod_stride, out_prod_stride;
    unsigned int res_prod_stride;
    unsigned long long half;   // pairs per table in the round being evaluated (N_j / 2)
    Fr* partials;              // [proof][block][npts] scratch
    unsigned int* counters;    // arrivals, groups done, work counters (kWorkCtrBase above); zero on entry, zero on exit
    unsigned int dynamic;      // 1: warp w of every CTA of a group works on the chunks (32 pairs) c = w mod 4 and takes them from that
                               // partition's counter (beyond a first, fixed one); 0: fixed stride.  The CTAs that share an SM do not
                               // progress at the same pace -- the schedulers favour the older warps -- so a fixed split leaves the SM
                               // half empty while its last CTA finishes (profiles/r02_resident_cta_spread.txt: c3 7.7 % faster); the
                               // proof does not depend on who folds which pair.  Needs all CTAs of the launch resident at once.
    Fr* result;                // [proof][res_stride]: npts Montgomery elements each
    unsigned int res_stride;
    unsigned int npts;         // evaluation points 0..npts-1 are wanted (<= D+1); SKIP1 kernels leave point 1 untouched
    volatile unsigned int* flag;  // optional: set to flag_value (system scope) after the results
    unsigned int flag_value;      // (0xffffffff is stored instead when the exchange timed out)
    XchArgs xch;                  // n_ranks > 1: exchange + reduce across GPUs before the flag is raised
};
// NB = proofs per launch.  The per-proof challenge data travels in the kernel parameters (constant bank): a
// single-proof launch carries 288 bytes of it, a batched one 18 KiB -- large parameter blocks make every
// launch slower, so the two cases are separate instantiations.
template <int NB>
struct RoundArgsT : RoundBase {
    Fr chal[NB];               // FOLD: challenge of the previous round, Montgomery form, per proof (degrees > 5)
    FoldTab tab[NB];           // FOLD: its shift table (fr.cuh mul_fixed_rows), per proof; read straight from the constant bank
};

#ifndef ZKSC_KARA_D2
#define ZKSC_KARA_D2 0
#endif
#ifndef ZKSC_KARA_D3
#define ZKSC_KARA_D3 0
#endif
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
// A per-thread accumulator that lives in shared memory, limb-major ([limb][thread]: conflict-free 32-bit accesses): the
// registers it frees matter where an instantiation sits at the register cap (experiment: ZKSC_SMEM_ACC).
template <int NL>
struct SmemAcc {
    uint32_t* p;   // &s_acc[slot][0][threadIdx.x]
};
template <int NL, int NX>
ZKSC_DEV void acc_add(SmemAcc<NL>& a, const uint32_t (&x)[NX]) {
    static_assert(NX <= NL, "");
    uint32_t v = a.p[0];
    a.p[0] = ptx::add_cc(v, x[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) {
        const uint32_t xi = (i < NX) ? x[i] : 0u;
        v = a.p[i * kThreads];
        a.p[i * kThreads] = (i < NL - 1) ? ptx::addc_cc(v, xi) : ptx::addc(v, xi);
    }
}
template <int D, class A>
ZKSC_DEV void accumulate_product(A& acc, const Fr (&f)[D]) {
    if constexpr (D == 1) {
        acc_add<9, 8>(acc, f[0].l);
    } else if constexpr (Lazy<D>::wide) {
        // ZKSC_KARA_D2 / ZKSC_KARA_D3: one level of Karatsuba in the 8 x 8 limb products of this degree (fr.cuh mul_wide_k)
        constexpr bool kKara = (D == 2 && ZKSC_KARA_D2) || (D == 3 && ZKSC_KARA_D3);
        Fr g = f[0];
#pragma unroll
        for (int k = 1; k < D - 1; k++) {
            if constexpr (kKara) g = (k == D - 2) ? fr_mul_lazy_k(g, f[k]) : fr_mul_k(g, f[k]);
            else g = (k == D - 2) ? fr_mul_lazy(g, f[k]) : fr_mul(g, f[k]);   // the one that feeds mul_wide may stay < 2r
        }
        uint32_t T[16];
        if constexpr (kKara) mul_wide_k(T, g, f[D - 1]);
        else mul_wide(T, g, f[D - 1]);      // last multiplication stays unreduced
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
// called by a whole warp, the sum in lane 0, the result in lane 0
template <int NL>
ZKSC_DEV Fr acc_finish_warp(const Acc<NL>& a) {
    if constexpr (NL == 9) return acc9_reduce(a);
    else return acc17_reduce_warp(a);
}

ZKSC_DEV Fr ld256_volatile(const Fr* p) {
    Fr v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]) : "l"(p) : "memory");
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7]) : "l"(p) : "memory");
    return v;
}
// Two 128-bit stores, for peer / IPC-exported / host-mapped destinations.  Measured on this pool's B200s
// (driver 580): a 256-bit st.global.v8.u32 (STG.E.ENL2.256) to memory that is exported through CUDA IPC -- local
// or peer side -- stored only its first 32-bit word; 128-bit stores behave.
ZKSC_DEV void st256_2x128(Fr* p, const Fr& v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.l[0]), "r"(v.l[1]), "r"(v.l[2]), "r"(v.l[3]) : "memory");
    asm volatile("st.global.v4.u32 [%0+16], {%1,%2,%3,%4};" ::"l"(p), "r"(v.l[4]), "r"(v.l[5]), "r"(v.l[6]), "r"(v.l[7]) : "memory");
}
ZKSC_DEV void st_release_sys(unsigned int* p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
ZKSC_DEV unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
ZKSC_DEV unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kXchTimeoutNs = 20ull * 1000 * 1000 * 1000;

// The whole (final) block: all-to-all of the partials through peer memory + modular sum.  Returns false on timeout.
static __device__ __noinline__ bool exchange_partials(const XchArgs& x) {
    __shared__ unsigned int s_ok;
    const unsigned int G = x.n_ranks, slot = (x.seq & 1u) * G;
    __threadfence();                                   // the partials were written by other blocks of this and earlier launches
    for (unsigned int i = threadIdx.x; i < x.n_elems; i += kThreads) {
        const Fr v = ld256_volatile(x.send + i);
        for (unsigned int g = 0; g < G; g++) st256_2x128(x.peer_data[g] + (size_t)(slot + x.rank) * x.cap + i, v);
    }
    if (threadIdx.x == 0) s_ok = 1u;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < G) {
        st_release_sys(x.peer_tags[threadIdx.x] + slot + x.rank, x.seq);
        const unsigned int* mine = x.peer_tags[x.rank] + slot + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(mine) != x.seq) {
            if (global_timer_ns() - t0 > kXchTimeoutNs) { s_ok = 0u; break; }
        }
    }
    __syncthreads();
    if (!s_ok) return false;
    __threadfence_system();
    const Fr* data = x.peer_data[x.rank] + (size_t)slot * x.cap;
    for (unsigned int i = threadIdx.x; i < x.n_elems; i += kThreads) {
        Fr s = ld256_volatile(data + i);
        for (unsigned int g = 1; g < G; g++) s = fr_add(s, ld256_volatile(data + (size_t)g * x.cap + i));
        st256_2x128(x.out + i, s);
    }
    __syncthreads();
    return true;
}

ZKSC_DEV Fr ld256_cg(const Fr* p) {     // L2 only: for data another SM wrote while this kernel was already running
    Fr v;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]), "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7])
                 : "l"(p)
                 : "memory");
    return v;
}
// The whole CTA (kThreads threads) sums, for each of NP evaluation points, n canonical elements  base[p * point_stride + i * elem_stride],
// i < n, into unreduced 9-limb accumulators  s_red[p][slice].  The points are summed side by side -- warp w takes (point, slice) =
// (w mod NP, w / NP), a slice being every kSlices-th element -- and ten loads per lane are issued before the first one is consumed:
// the last block of a round walks up to 148 x 4 per-block partials per point, and one point after the other with one load per
// dependent iteration made that walk 3-7 us long.
template <int NP>
struct SumShape {
    static constexpr int kSlices = (NP <= kWarps) ? kWarps / NP : 1;     // warps per point
};
template <int NP>
ZKSC_DEV void cta_sum_points(const Fr* base, size_t point_stride, size_t elem_stride, unsigned int n, Acc<9> (*s_red)[kWarps]) {
    constexpr int KB = 10, SL = SumShape<NP>::kSlices;
    const int lane = threadIdx.x & 31;
    for (int task = threadIdx.x >> 5; task < NP * SL; task += kWarps) {
        const int p = task % NP, j = task / NP;
        const Fr* src = base + (size_t)p * point_stride;
        Acc<9> a;
        acc_zero(a);
        for (unsigned int b0 = j * 32; b0 < n; b0 += KB * SL * 32) {
            Fr v[KB];
#pragma unroll
            for (int k = 0; k < KB; k++) {
                const unsigned int c = b0 + k * SL * 32 + lane;
                v[k] = (c < n) ? ld256_cg(src + (size_t)c * elem_stride) : fr_zero();
            }
#pragma unroll
            for (int k = 0; k < KB; k++) acc_add<9, 8>(a, v[k].l);
        }
        acc_warp_reduce(a);
        if (lane == 0) s_red[p][j] = a;
    }
}
// ... and one warp finishes a point: the canonical total of what cta_sum_points left in s_red[p] (call after a __syncthreads)
template <int NP>
ZKSC_DEV Fr warp_finish_sum(const Acc<9>* s_red_p) {
    const int lane = threadIdx.x & 31;
    Acc<9> a;
    if (lane < SumShape<NP>::kSlices) a = s_red_p[lane];
    else acc_zero(a);
    acc_warp_reduce(a);
    return acc9_reduce(a);      // meaningful in lane 0
}

// Block-level reduction of NP accumulators, publication of the block partial, and -- in the last
// block of each proof to arrive -- the final cross-block sum.
// Accumulator slot s holds evaluation point s, or with SKIP1 point (s == 0 ? 0 : s + 1).
template <int NL, int NP, bool SKIP1>
ZKSC_DEV void reduce_and_publish(Acc<NL> (&acc)[NP], const RoundBase& args, int npts) {
    auto point_of = [](int slot) { return (SKIP1 && slot > 0) ? slot + 1 : slot; };
    __shared__ Acc<NL> s_warp[kWarps][NP];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int proof = blockIdx.y;
    const unsigned int group = blockIdx.z * gridDim.y + proof;      // (product, proof) of this launch
    const unsigned int n_groups = gridDim.y * gridDim.z;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        acc_warp_reduce(acc[p]);
        if (lane == 0) s_warp[warp][p] = acc[p];
    }
    __syncthreads();
    Fr* my_partials = args.partials + ((size_t)group * gridDim.x + blockIdx.x) * NP;
    if (warp == 0) {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            Acc<NL> a;
            if (lane < kWarps) a = s_warp[lane][p];
            else acc_zero(a);
            acc_warp_reduce(a);
            const Fr v = acc_finish_warp<NL>(a);
            if (lane == 0 && point_of(p) < npts) st256(my_partials + p, v);
        }
        if (lane == 0) {
            __threadfence();
            unsigned int done = atomicAdd(args.counters + group, 1u);
            s_last = (done == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block of this proof: sum the per-block partials (canonical Montgomery elements)
    const Fr* all = args.partials + (size_t)group * gridDim.x * NP;
    __shared__ Acc<9> s_red[NP][kWarps];
    cta_sum_points<NP>(all, 1, NP, gridDim.x, s_red);
    __syncthreads();
    for (int p = warp; p < NP; p += kWarps) {
        const Fr v = warp_finish_sum<NP>(s_red[p]);
        if (lane == 0 && point_of(p) < npts)
            st256_2x128(args.result + (size_t)proof * args.res_stride + (size_t)blockIdx.z * args.res_prod_stride + point_of(p), v);   // may be host-mapped
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        args.counters[group] = 0;
        if (args.dynamic)
            for (int w = 0; w < kWarps; w++) args.counters[kWorkCtrBase + (group * kWarps + w) * kWorkCtrWords] = 0;
        bool final_block = false;
        if (args.flag) {
            __threadfence_system();
            // one flag word per proof would be wasteful: the groups bump a shared word via the counter slot behind theirs
            unsigned int fin = atomicAdd(args.counters + n_groups, 1u);
            if (fin == n_groups - 1) {
                args.counters[n_groups] = 0;
                final_block = true;
            }
        }
        s_last = final_block;
    }
    __syncthreads();
    if (!s_last) return;
    // the block that finished the launch's last proof: exchange across GPUs if asked, then raise the flag
    unsigned int flag_value = args.flag_value;
    if (args.xch.n_ranks > 1 && !exchange_partials(args.xch)) flag_value = 0xffffffffu;
    if (threadIdx.x == 0) {
        __threadfence_system();
        *args.flag = flag_value;
    }
}

// acc[] += the products at every evaluation point of one pair (a = even half, b = odd half of each factor)
template <int D, bool SKIP1, class A, int NP>
ZKSC_DEV void accumulate_points(A (&acc)[NP], Fr (&a)[D], Fr (&b)[D], int npts) {
    // evaluation points 0 and 1 are the two halves themselves
    accumulate_product<D>(acc[0], a);
    if (!SKIP1 && npts > 1) accumulate_product<D>(acc[1], b);
    if constexpr (D == 2) {
        // Degree 2: the third value is the LEADING COEFFICIENT  h(inf) = sum (b_0 - a_0)(b_1 - a_1)  instead of h(2): both factors
        // enter mul_wide, so the differences may stay unreduced (b - a + r in (0, 2r): 16 ALU instructions per factor instead of
        // 33 for a canonical difference plus the walk to t = 2).  The host turns it into h(2) = 2 h(1) - h(0) + 2 h(inf)
        // (zksc.cu finish_round) -- exact field arithmetic, identical canonical value.
        if (npts > 2) {
            Fr dl[2];
#pragma unroll
            for (int k = 0; k < 2; k++) dl[k] = fr_sub_lazy(b[k], a[k]);
            accumulate_product<D>(acc[SKIP1 ? 1 : 2], dl);
        }
    } else if constexpr (D == 3) {
        // Degree 3: the third and fourth values are h(-1) and the leading coefficient h(inf) instead of h(2), h(3):
        //   f_k(-1) = a_k - delta_k,  f_k(inf) = delta_k,  delta_k = b_k - a_k
        // -- 132 ALU instructions per pair for the operands instead of 191 for the walk to t = 2, 3 (the first factor must be
        // canonical for fr_mul_lazy, the other two may stay unreduced).  The host solves for h(2), h(3) (zksc.cu finish_round).
        if (npts > 2) {
            Fr dl[3], m[3];
#pragma unroll
            for (int k = 0; k < 3; k++) dl[k] = fr_sub(b[k], a[k]);
            m[0] = fr_sub(a[0], dl[0]);
            m[1] = fr_sub_lazy(a[1], dl[1]);
            m[2] = fr_sub_lazy(a[2], dl[2]);
            accumulate_product<D>(acc[SKIP1 ? 1 : 2], m);
            if (npts > 3) accumulate_product<D>(acc[SKIP1 ? 2 : 3], dl);
        }
    } else if (D >= 2 && npts > 2) {
        // f_k(t) = a_k + t (b_k - a_k): walk t = 2..D by repeated addition of the difference
        Fr delta[D];
#pragma unroll
        for (int k = 0; k < D; k++) delta[k] = fr_sub(b[k], a[k]);
#pragma unroll
        for (int p = 2; p <= D; p++) {
            // The last point's values are not walked any further, so they need not be canonical: b + delta < 2r goes straight
            // into the product -- every factor but the first only ever meets a canonical partner in fr_mul (x * y < r 2^256), the
            // last one enters mul_wide (any 256-bit value); with D = 2 both factors enter mul_wide.
#pragma unroll
            for (int k = 0; k < D; k++) {
                const bool lazy = (p == D) && (k >= 1 || D == 2);
                b[k] = lazy ? fr_add_lazy(b[k], delta[k]) : fr_add(b[k], delta[k]);
            }
            if (p < npts) accumulate_product<D>(acc[SKIP1 ? p - 1 : p], b);
        }
    }
}

// FOLD = false : evaluate the round polynomial of the tables as they are (first round).
// FOLD = true  : bind the previous challenge (in -> out), then evaluate the next round on the result.
#ifndef ZKSC_ROUND_MINB
#define ZKSC_ROUND_MINB 4   // caps the round kernels at 128 registers; measured 1-3 % faster than ptxas' own choice (profiles/r01_variants_v4.txt)
#endif
#ifndef ZKSC_FOLD_LOADALL
#define ZKSC_FOLD_LOADALL 1
#endif
#ifndef ZKSC_PREFETCH
#define ZKSC_PREFETCH 0
#endif
ZKSC_DEV void prefetch_line(const void* p) {
#if ZKSC_PREFETCH == 2
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
template <int D, bool FOLD, bool SKIP1, int NB>
__global__ void __launch_bounds__(kThreads, ZKSC_ROUND_MINB) round_kernel(const __grid_constant__ RoundArgsT<NB> args) {
    constexpr int NL = Lazy<D>::NL;
#if defined(ZKSC_FOLD_TWO_CONDSUB)
    constexpr bool kSemi = false;
#elif defined(ZKSC_SEMI_ALL_NB1)
    constexpr bool kSemi = (NB == 1);             // experiment
#else
    constexpr bool kSemi = (D == 2 && NB == 1);   // see fr_fold_tab
#endif
    constexpr int NP = SKIP1 ? D : D + 1;   // accumulators
    const int proof = (NB == 1) ? 0 : blockIdx.y;
    const Fr* in = args.in + (size_t)proof * args.in_proof_stride + (size_t)blockIdx.z * args.in_prod_stride;
    Fr* out = args.out + (size_t)proof * args.out_proof_stride + (size_t)blockIdx.z * args.out_prod_stride;
    const unsigned long long half = args.half;
    const int npts = args.npts;

    Fr r;
    if constexpr (FOLD && D > 5) r = args.chal[proof];

#ifndef ZKSC_SMEM_ACC
#define ZKSC_SMEM_ACC 3
#endif
    // Accumulators in shared memory for the fold rounds of degree ZKSC_SMEM_ACC (0 = none).  The d = 3 fold keeps 3 x 17 limbs of
    // sums next to 6 folded entries and sits at the 128-register cap with 144 bytes of spills; with the sums in shared memory
    // (27-35 KB per CTA) it needs 120 registers, spills nothing and runs 1.2 % faster (profiles/r01_variants_v5.txt).
    constexpr bool kSmemAcc = (ZKSC_SMEM_ACC != 0) && (D == ZKSC_SMEM_ACC) && FOLD && Lazy<D>::wide;
    __shared__ uint32_t s_acc[kSmemAcc ? NP * NL * kThreads : 1];
    SmemAcc<NL> sacc[NP];
    Acc<NL> acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        acc_zero(acc[p]);
        if constexpr (kSmemAcc) {
            sacc[p].p = s_acc + (size_t)p * NL * kThreads + threadIdx.x;
#pragma unroll
            for (int i = 0; i < NL; i++) sacc[p].p[i * kThreads] = 0u;
        }
    }

    // chunk c = pairs [32 c, 32 c + 32): a warp's first chunk is its index in the group, the following ones come at a fixed stride
    // or from the group's counter (asked for before the current chunk is worked on: the atomic's latency is hidden)
    const unsigned int n_chunks = (unsigned int)((half + 31) >> 5);
    unsigned int* work_ctr = args.counters + kWorkCtrBase + ((blockIdx.z * gridDim.y + proof) * kWarps + (threadIdx.x >> 5)) * kWorkCtrWords;
    const bool dynamic = args.dynamic != 0u;
    for (unsigned int i = blockIdx.x;;) {                    // i-th chunk of this warp's partition
        const unsigned int c = i * kWarps + (threadIdx.x >> 5);
        if (c >= n_chunks) break;
        unsigned int nxt = i + gridDim.x;
        if (dynamic && (threadIdx.x & 31) == 0) nxt = atomicAdd(work_ctr, 1u) + gridDim.x;
        const unsigned long long x = ((unsigned long long)c << 5) + (threadIdx.x & 31);
        i = dynamic ? __shfl_sync(0xffffffffu, nxt, 0) : nxt;
        if (x >= half) continue;
        Fr a[D], b[D];
        if constexpr (FOLD && D == 2 && ZKSC_FOLD_LOADALL) {
            // all eight entries of the pair in flight at once: one exposed memory latency per iteration instead of two
            const Fr* t0 = in;
            const Fr* t1 = in + args.in_tab_stride;
            Fr p0 = ld256(t0 + x), p1 = ld256(t0 + x + 2 * half), q0 = ld256(t0 + x + half), q1 = ld256(t0 + x + 3 * half);
            Fr r0 = ld256(t1 + x), r1 = ld256(t1 + x + 2 * half), s0 = ld256(t1 + x + half), s1 = ld256(t1 + x + 3 * half);
            a[0] = fr_fold_tab<kSemi>(p0, p1, args.tab[proof]);
            b[0] = fr_fold_tab<kSemi>(q0, q1, args.tab[proof]);
            st256(out + x, a[0]);
            st256(out + x + half, b[0]);
            a[1] = fr_fold_tab<kSemi>(r0, r1, args.tab[proof]);
            b[1] = fr_fold_tab<kSemi>(s0, s1, args.tab[proof]);
            st256(out + args.out_tab_stride + x, a[1]);
            st256(out + args.out_tab_stride + x + half, b[1]);
        } else
#pragma unroll
        for (int k = 0; k < D; k++) {
            const Fr* t = in + (size_t)k * args.in_tab_stride;
            if constexpr (FOLD) {
                // T_{j-1} has 4*half entries; its pairs are (y, y + 2*half)
                Fr p0 = ld256(t + x), p1 = ld256(t + x + 2 * half);
                Fr q0 = ld256(t + x + half), q1 = ld256(t + x + 3 * half);
                if constexpr (D <= 5) {
                    a[k] = fr_fold_tab<kSemi>(p0, p1, args.tab[proof]);
                    b[k] = fr_fold_tab<kSemi>(q0, q1, args.tab[proof]);
                } else {
                    a[k] = fr_fold_d<D>(p0, p1, r);
                    b[k] = fr_fold_d<D>(q0, q1, r);
                }
                Fr* o = out + (size_t)k * args.out_tab_stride;
                st256(o + x, a[k]);
                st256(o + x + half, b[k]);
            } else {
                a[k] = ld256_stream(t + x);
                b[k] = ld256_stream(t + x + half);
            }
        }
        if constexpr (kSmemAcc) accumulate_points<D, SKIP1>(sacc, a, b, npts);
        else accumulate_points<D, SKIP1>(acc, a, b, npts);
    }
    if constexpr (kSmemAcc) {
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int i = 0; i < NL; i++) acc[p].l[i] = sacc[p].p[i * kThreads];
    }
    reduce_and_publish<NL, NP, SKIP1>(acc, args, npts);
}

// ---- TMA-staged variant ---------------------------------------------------------------------------------
// Same arithmetic, different data movement.  The kernel above issues its global loads at the top of every
// iteration and then waits for them; with ~450-1000 multiplier-pipe instructions per pair only four warps
// fit per scheduler, too few to cover HBM latency that way.  Here every warp owns one shared-memory slot
// (32 pairs: 4 D segments of 1 KiB with FOLD, 2 D without) and one mbarrier; lane 0 fills the slot with
// cp.async.bulk (the TMA engine, UBLKCP in SASS) for the warp's NEXT tile as soon as the current tile has
// been copied from the slot into registers, so the whole compute phase of a tile overlaps the fetch of the
// next one.  No producer warp, no cross-warp synchronisation: the slot is released by the warp that owns it.
// The loads carry an L2 evict_first policy (T_{j-1} is dead after this pass) so that the folded tables
// just written stay in L2 for the next round when they fit.
// Requires half % 32 == 0 (the host falls back to round_kernel for the tiny rounds).
ZKSC_DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
ZKSC_DEV void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
ZKSC_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
ZKSC_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ZKSC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ZKSC_DONE_%=;\n"
        "bra ZKSC_WAIT_%=;\n"
        "ZKSC_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
ZKSC_DEV unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
ZKSC_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
// plain C++ loads (2 x LDS.128): ordered by the compiler against the mbarrier wait / __syncwarp around them,
// free to be scheduled between the two
ZKSC_DEV Fr lds256(const unsigned char* p) {
    const uint4 lo = *reinterpret_cast<const uint4*>(p), hi = *reinterpret_cast<const uint4*>(p + 16);
    Fr v;
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
    v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
    return v;
}

#ifndef ZKSC_TMA_MINB
#define ZKSC_TMA_MINB(D) ((D) <= 2 ? 4 : 3)
#endif
constexpr int kTilePairs = 32;                       // pairs per warp tile: one per lane
constexpr int kSegBytes = kTilePairs * 32;           // one table segment of a tile
template <int D, bool FOLD>
__host__ __device__ constexpr int tma_slot_bytes() { return (FOLD ? 4 : 2) * D * kSegBytes; }

template <int D, bool FOLD, bool SKIP1, int NB>
__global__ void __launch_bounds__(kThreads, ZKSC_TMA_MINB(D)) round_tma_kernel(const __grid_constant__ RoundArgsT<NB> args) {
    static_assert(D <= 5, "the staged kernel uses the table fold");
    constexpr int NL = Lazy<D>::NL;
    constexpr int NP = SKIP1 ? D : D + 1;
    constexpr int NSEG = FOLD ? 4 : 2;
    constexpr int SLOT = tma_slot_bytes<D, FOLD>();
    extern __shared__ __align__(128) unsigned char zksc_slots[];
    __shared__ __align__(8) unsigned long long s_bar[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int proof = (NB == 1) ? 0 : blockIdx.y;
    const Fr* in = args.in + (size_t)proof * args.in_proof_stride + (size_t)blockIdx.z * args.in_prod_stride;
    Fr* out = args.out + (size_t)proof * args.out_proof_stride + (size_t)blockIdx.z * args.out_prod_stride;
    const unsigned long long half = args.half;
    const int npts = args.npts;
    const unsigned char* my_slot = zksc_slots + warp * SLOT;
    const uint32_t slot = smem_addr(my_slot);
    const uint32_t bar = smem_addr(&s_bar[warp]);
    // 32-bit tile counters: half / 32 < 2^32 for any table that fits in memory
    const unsigned int n_tiles = (unsigned int)(half / kTilePairs);
    const unsigned int tile_stride = gridDim.x * kWarps;
    unsigned int tile = blockIdx.x * kWarps + warp;
    // a warp's first two tiles are fixed (the second is fetched while the first is worked on); the following ones come at a fixed
    // stride or, with args.dynamic, from the counter of the warp's partition (tiles = warp mod 4) -- asked for one tile ahead of the
    // fetch (see RoundBase::dynamic)
    unsigned int next = tile + tile_stride;
    unsigned int* work_ctr = args.counters + kWorkCtrBase + ((blockIdx.z * gridDim.y + proof) * kWarps + warp) * kWorkCtrWords;
    const bool dynamic = args.dynamic != 0u;

    auto issue = [&](unsigned int t) {               // lane 0 only
        const unsigned long long x0 = (unsigned long long)t * kTilePairs;
        const unsigned long long policy = l2_evict_first_policy();
        mbar_expect_tx(bar, SLOT);
#pragma unroll
        for (int k = 0; k < D; k++)
#pragma unroll
            for (int sgm = 0; sgm < NSEG; sgm++)
                bulk_g2s(slot + (k * NSEG + sgm) * kSegBytes, in + (size_t)k * args.in_tab_stride + x0 + sgm * half, kSegBytes, bar, policy);
    };
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (tile < n_tiles) issue(tile);
    }
    __syncwarp();

    Acc<NL> acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) acc_zero(acc[p]);

    uint32_t phase = 0;
    while (tile < n_tiles) {
        const unsigned long long x = (unsigned long long)tile * kTilePairs + lane;
        unsigned int fut = next + tile_stride;
        if (dynamic && lane == 0) fut = (atomicAdd(work_ctr, 1u) + 2 * gridDim.x) * kWarps + warp;
        mbar_wait(bar, phase);
        phase ^= 1;
        Fr a[D], b[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            // operands are copied out of the slot just before they are used (two entries at a time with FOLD),
            // which keeps the live registers low; the slot is released after the last copy
            const unsigned char* tab = my_slot + k * NSEG * kSegBytes + lane * 32;
            if constexpr (FOLD) {
                // segments: x, x + half, x + 2 half, x + 3 half of T_{j-1}; its pairs are (y, y + 2 half)
                Fr* o = out + (size_t)k * args.out_tab_stride;
                {
                    const Fr p0 = lds256(tab), p1 = lds256(tab + 2 * kSegBytes);
                    a[k] = fr_fold_tab(p0, p1, args.tab[proof]);
                    st256(o + x, a[k]);
                }
                const Fr q0 = lds256(tab + kSegBytes), q1 = lds256(tab + 3 * kSegBytes);
                if (k == D - 1) {
                    __syncwarp();                     // every lane has its last operands: the slot is free
                    if (lane == 0 && next < n_tiles) issue(next);
                }
                b[k] = fr_fold_tab(q0, q1, args.tab[proof]);
                st256(o + x + half, b[k]);
            } else {
                a[k] = lds256(tab);
                b[k] = lds256(tab + kSegBytes);
                if (k == D - 1) {
                    __syncwarp();
                    if (lane == 0 && next < n_tiles) issue(next);
                }
            }
        }
        accumulate_points<D, SKIP1>(acc, a, b, npts);
        tile = next;
        next = dynamic ? __shfl_sync(0xffffffffu, fut, 0) : fut;
    }
    reduce_and_publish<NL, NP, SKIP1>(acc, args, npts);
}

}  // namespace zksc
