// K4 -- the persistent tail kernel: every remaining round of a proof once its tables are small.
//
// From the round whose tables have at most 2 * kTailPairs entries on, a round is latency, not bandwidth:
// a launch (~4 us), a cold instruction cache, the cross-block reduction through global memory and a
// completion word cost ~25 us per round against < 2 us of arithmetic.  This kernel is launched ONCE per
// proof tail and stays resident: one CTA per (proof, product) loops over the rounds
//     wait for the challenge of the previous round (mailbox in pinned host memory, polled over PCIe)
//     fold every table in place with it and accumulate the next round's evaluations   (same arithmetic
//       as round_kernel<D, FOLD = true, SKIP1 = true>: kernels.cuh)
//     reduce inside the CTA (warp shuffles + shared memory; no global partials, no atomics)
//     publish the evaluations into pinned host memory
// The Fiat-Shamir transcript stays on the host (north_star): the host thread polls the published evaluations,
// absorbs the round polynomial, derives the challenge and posts its fold table to the mailbox.
//
// Both directions use self-validating 8-byte units {payload word, sequence number}: an aligned 8-byte store is
// atomic on the host and on the device, so a reader that sees the expected sequence number in EVERY unit has the
// whole message -- no fences, no ordering assumptions about PCIe.
//
// Replaces, per round, multi_composed_sumcheck.rs:81-89 + :103-105 of the reference, like the round kernels.
#pragma once
#include "kernels.cuh"

namespace zksc {

constexpr int kTailThreads = 256;
constexpr int kTailWarps = kTailThreads / 32;
constexpr unsigned long long kTailPairs = 1024;     // the tail starts at rounds with at most this many pairs per table
constexpr int kTailMaxDegree = 5;                   // table fold (fr.cuh mul_fixed_rows) degrees
constexpr int kTailMaxProducts = 8;                 // == ZKSC_MAX_PRODUCTS
constexpr int kMailUnits = 64;                      // FoldTab words per proof and round
constexpr unsigned int kTailAbort = 0xffffffffu;    // sequence number that tells the kernel to leave
constexpr unsigned int kTailTimeout = 0xfffffffeu;  // published by the kernel when no challenge arrived in time
constexpr unsigned long long kTailTimeoutNs = 10ull * 1000 * 1000 * 1000;

struct TailArgs {
    const Fr* in;              // tables before the first fold of the tail (table 0 of proof 0)
    Fr* out;                   // folded tables (may alias `in`)
    unsigned long long in_tab_stride, in_proof_stride, out_tab_stride, out_proof_stride;   // elements
    unsigned long long half;   // pairs per table of the first round evaluated here
    unsigned int n_rounds;     // rounds to run (half, half/2, ..., 1 pairs when run to the end)
    unsigned int seq0;         // sequence number of the first round's challenge and result
    unsigned int n_products, n_evals;          // products per proof; sum of (degree + 1)
    unsigned int deg[kTailMaxProducts], koff[kTailMaxProducts], eoff[kTailMaxProducts];
    const uint2* mail;         // [proof][kMailUnits] host-mapped: {fold-table word, seq}
    uint2* results;            // [proof][n_evals][8] host-mapped: {limb, seq}
};

#ifdef ZKSC_TAIL_IMPL   // the kernel itself: tail_inst.cu only (zksc.cu needs just the declarations above)
ZKSC_DEV uint4 ld_volatile_v4(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
ZKSC_DEV void st_volatile_v2(void* p, unsigned int a, unsigned int b) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

// Warp 0 polls the proof's mailbox (two units per lane = one 16-byte read) until every unit carries `seq`.
// Returns 0 = challenge in `tab`, 1 = abort requested, 2 = timed out.   Called by warp 0 only.
ZKSC_DEV int tail_wait_mail(const uint2* mail, unsigned int seq, FoldTab& tab) {
    const int lane = threadIdx.x & 31;
    const unsigned long long t0 = global_timer_ns();
    for (unsigned int spins = 1;; spins++) {
        const uint4 v = ld_volatile_v4(mail + 2 * lane);
        const bool ok = (v.y == seq) && (v.w == seq);
        const bool abort = (lane == 0) && (v.y == kTailAbort);
        if (__any_sync(0xffffffffu, abort)) return 1;
        if (__all_sync(0xffffffffu, ok)) {
            uint32_t* w = &tab.w[0][0];
            w[2 * lane] = v.x;
            w[2 * lane + 1] = v.z;
            return 0;
        }
        if ((spins & 0x3ff) == 0 && global_timer_ns() - t0 > kTailTimeoutNs) return 2;
    }
}

// One (proof, product) CTA of degree D: all rounds.
template <int D>
ZKSC_DEV void tail_body(const TailArgs& args, const int proof, const int product, FoldTab& s_tab, int& s_state) {
    constexpr int NL = Lazy<D>::NL;
    constexpr int NP = D;                       // point 1 is derived on the host (SKIP1)
    __shared__ Acc<NL> s_warp[kTailWarps][NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int koff = args.koff[product];
    const Fr* in = args.in + (size_t)proof * args.in_proof_stride + (size_t)koff * args.in_tab_stride;
    unsigned long long in_stride = args.in_tab_stride;
    Fr* const out = args.out + (size_t)proof * args.out_proof_stride + (size_t)koff * args.out_tab_stride;
    const uint2* mail = args.mail + (size_t)proof * kMailUnits;
    uint2* res = args.results + ((size_t)proof * args.n_evals + args.eoff[product]) * 8;
    unsigned long long half = args.half;

    for (unsigned int round = 0; round < args.n_rounds; round++, half >>= 1) {
        const unsigned int seq = args.seq0 + round;
        if (warp == 0) {
            const int st = tail_wait_mail(mail, seq, s_tab);
            if (lane == 0) s_state = st;
        }
        __syncthreads();
        if (s_state != 0) {
            if (s_state == 2 && threadIdx.x < 8 * (D + 1)) st_volatile_v2(res + threadIdx.x, 0u, kTailTimeout);
            return;
        }
        Acc<NL> acc[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) acc_zero(acc[p]);
        for (unsigned long long x = threadIdx.x; x < half; x += kTailThreads) {
            Fr a[D], b[D];
#pragma unroll
            for (int k = 0; k < D; k++) {
                const Fr* t = in + (size_t)k * in_stride;
                const Fr p0 = ld256(t + x), p1 = ld256(t + x + 2 * half);
                const Fr q0 = ld256(t + x + half), q1 = ld256(t + x + 3 * half);
                a[k] = fr_fold_tab(p0, p1, s_tab);
                b[k] = fr_fold_tab(q0, q1, s_tab);
                Fr* o = out + (size_t)k * args.out_tab_stride;
                st256(o + x, a[k]);
                st256(o + x + half, b[k]);
            }
            accumulate_points<D, true>(acc, a, b, D + 1);
        }
        // reduce inside the CTA
#pragma unroll
        for (int p = 0; p < NP; p++) {
            acc_warp_reduce(acc[p]);
            if (lane == 0) s_warp[warp][p] = acc[p];
        }
        __syncthreads();      // also: this round's stores to `out` are visible to the whole CTA, s_tab may be rewritten
        if (warp == 0) {
#pragma unroll
            for (int p = 0; p < NP; p++) {
                Acc<NL> a;
                if (lane < kTailWarps) a = s_warp[lane][p];
                else acc_zero(a);
                acc_warp_reduce(a);
                Fr v = acc_finish<NL>(a);                          // lane 0 holds the total
                const int point = (p == 0) ? 0 : p + 1;
#pragma unroll
                for (int l = 0; l < 8; l++) {
                    const uint32_t limb = __shfl_sync(0xffffffffu, v.l[l], 0);
                    if (lane == l) st_volatile_v2(res + point * 8 + l, limb, seq);
                }
            }
        }
        in = out;
        in_stride = args.out_tab_stride;
    }
}

__global__ void __launch_bounds__(kTailThreads, 1) tail_kernel(const __grid_constant__ TailArgs args) {
    __shared__ FoldTab s_tab;
    __shared__ int s_state;
    const int proof = blockIdx.x, product = blockIdx.y;
    switch (args.deg[product]) {
        case 1: tail_body<1>(args, proof, product, s_tab, s_state); break;
        case 2: tail_body<2>(args, proof, product, s_tab, s_state); break;
        case 3: tail_body<3>(args, proof, product, s_tab, s_state); break;
        case 4: tail_body<4>(args, proof, product, s_tab, s_state); break;
        case 5: tail_body<5>(args, proof, product, s_tab, s_state); break;
        default: break;
    }
}

#endif  // ZKSC_TAIL_IMPL

}  // namespace zksc
