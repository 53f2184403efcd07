// K4 -- the resident rounds kernel: every remaining round of a proof once a round is latency, not bandwidth.
//
// From the round whose arithmetic drops under kTailWorkPerCta limb products per resident CTA on, an ordinary
// round costs a launch (~4 us), a cold start, a cross-block reduction through global atomics and a completion
// word: ~25 us per round against a few us of arithmetic (and ~30 us on a sharded context, where the partial
// evaluations also cross NVLink).  This kernel is launched ONCE for all those rounds and stays resident: a
// group of up to `n_ctas` CTAs per (proof, product) loops over the rounds
//     CTA 0 waits for the previous round's challenge (mailbox in pinned host memory, polled over PCIe) and
//       relays it to the other CTAs of its group through a mailbox in HBM
//     every CTA folds its share of every table in place with it and accumulates the next round's
//       evaluations   (the arithmetic of round_kernel<D, FOLD = true, SKIP1 = true>: kernels.cuh)
//     CTA-level reduction (warp shuffles + shared memory); the other CTAs hand their sums to CTA 0 through HBM
//     sharded contexts: CTA 0 stores the group's sums into every rank's exchange buffer (peer memory over
//       NVLink), waits for the other ranks' sums and adds them mod r
//     CTA 0 publishes the evaluations into pinned host memory
// CTAs whose share of the table is empty leave (tables only shrink), so the last rounds run in CTA 0 alone with
// no hand-over at all.  The Fiat-Shamir transcript stays on the host (north_star): the host thread polls the
// published evaluations, absorbs the round polynomial, derives the challenge and posts its fold table.
//
// Every message -- host mailbox, relay, CTA sums, cross-GPU sums, published evaluations -- is made of
// self-validating 64-bit units  payload | (sequence number << 32), written and read with single 64-bit accesses
// (atomic on host and device).  A reader that finds the expected sequence number in EVERY unit has the whole
// message: no flags, no ordering assumptions about PCIe or NVLink.  Table data written by one CTA and read by
// another in the next round is ordered by fences around those messages (release: fence, then the CTA's sum
// units; acquire: relay units, then fence) and read with L1-bypassing loads.
//
// Replaces, per round, multi_composed_sumcheck.rs:81-89 + :103-105 of the reference, like the round kernels.
#pragma once
#include "kernels.cuh"

namespace zksc {

constexpr int kTailThreads = 256;
constexpr int kTailWarps = kTailThreads / 32;
constexpr unsigned long long kTailWorkPerCta = 920000;  // start when limb products per round <= this x CTAs of a group (see tail_eligible)
constexpr int kTailMaxCtas = 128;                   // CTAs per (proof, product) group
constexpr int kTailMaxDegree = 5;                   // table fold (fr.cuh mul_fixed_rows) degrees
constexpr int kTailMaxProducts = 8;                 // == ZKSC_MAX_PRODUCTS
constexpr int kMailUnits = 64;                      // FoldTab words per proof and round
constexpr unsigned int kTailAbort = 0xffffffffu;    // sequence number that tells the kernel to leave
constexpr unsigned int kTailTimeout = 0xfffffffeu;  // published when no challenge arrived in time (nothing folded: recoverable)
constexpr unsigned int kTailFailed = 0xfffffffdu;   // published when a CTA or a peer GPU went missing in the middle of a round
constexpr unsigned long long kTailTimeoutNs = 2ull * 1000 * 1000 * 1000;    // CTA 0 waiting for the host; everything else waits a multiple

struct TailArgs {
    const Fr* in;              // tables before the first fold of the tail (table 0 of proof 0)
    Fr* out;                   // folded tables (may alias `in`)
    unsigned long long in_tab_stride, in_proof_stride, out_tab_stride, out_proof_stride;   // elements
    unsigned long long half;   // pairs per table of the first round evaluated here
    unsigned int n_rounds;     // rounds to run (half, half/2, ..., 1 pairs when run to the end)
    unsigned int seq0;         // sequence number of the first round's messages
    unsigned int n_products, n_evals;          // products per proof; sum of (degree + 1)
    unsigned int deg[kTailMaxProducts], koff[kTailMaxProducts], eoff[kTailMaxProducts];
    const unsigned long long* mail;   // [proof][kMailUnits]                 host-mapped: fold-table word | seq << 32
    unsigned long long* results;      // [proof][n_evals][8]                 host-mapped: limb | seq << 32
    unsigned long long* relay;        // [group][kMailUnits]                 HBM   (group = proof * n_products + product)
    unsigned long long* sums;         // [group][n_ctas][kTailMaxDegree][8]  HBM
    // cross-GPU exchange (n_ranks > 1): every rank's unit buffer [2][n_ranks][xch_cap][8]
    unsigned long long* peer_units[kMaxRanks];
    unsigned int n_ranks, rank, xch_cap;
};

#ifdef ZKSC_TAIL_IMPL   // the kernel itself: tail_inst.cu only (zksc.cu needs just the declarations above)
ZKSC_DEV void ld_units2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
ZKSC_DEV void st_unit(unsigned long long* p, unsigned int payload, unsigned int seq) {
    const unsigned long long v = (unsigned long long)payload | ((unsigned long long)seq << 32);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
ZKSC_DEV void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
ZKSC_DEV Fr ld256_cg(const Fr* p) {     // L2 only: the entry may have been written by another SM in the previous round
    Fr v;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]), "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7])
                 : "l"(p)
                 : "memory");
    return v;
}

// Warp-level: poll a 64-unit mailbox (two units per lane) until every unit carries `seq`; the fold table lands in `tab`.
// Returns 0 = ok, 1 = abort requested (unit 0 carries kTailAbort), 2 = timed out.
ZKSC_DEV int tail_wait_mail(const unsigned long long* mail, unsigned int seq, FoldTab& tab, unsigned long long timeout_ns) {
    const int lane = threadIdx.x & 31;
    const unsigned long long t0 = global_timer_ns();
    for (unsigned int spins = 1;; spins++) {
        unsigned long long a, b;
        ld_units2(mail + 2 * lane, a, b);
        const bool ok = ((unsigned int)(a >> 32) == seq) && ((unsigned int)(b >> 32) == seq);
        const bool abort = (lane == 0) && ((unsigned int)(a >> 32) == kTailAbort);
        if (__any_sync(0xffffffffu, abort)) return 1;
        if (__all_sync(0xffffffffu, ok)) {
            uint32_t* w = &tab.w[0][0];
            w[2 * lane] = (unsigned int)a;
            w[2 * lane + 1] = (unsigned int)b;
            return 0;
        }
        if ((spins & 0x3ff) == 0 && global_timer_ns() - t0 > timeout_ns) return 2;
    }
}

// One lane: read an element published as 8 units; spins until all carry `seq`.  false on timeout.
ZKSC_DEV bool tail_read_elem(const unsigned long long* u, unsigned int seq, Fr& out) {
    const unsigned long long t0 = global_timer_ns();
    for (unsigned int spins = 1;; spins++) {
        unsigned long long v[8];
        ld_units2(u, v[0], v[1]);
        ld_units2(u + 2, v[2], v[3]);
        ld_units2(u + 4, v[4], v[5]);
        ld_units2(u + 6, v[6], v[7]);
        bool ok = true;
#pragma unroll
        for (int l = 0; l < 8; l++) ok = ok && ((unsigned int)(v[l] >> 32) == seq);
        if (ok) {
#pragma unroll
            for (int l = 0; l < 8; l++) out.l[l] = (unsigned int)v[l];
            return true;
        }
        if ((spins & 0x3ff) == 0 && global_timer_ns() - t0 > 10 * kTailTimeoutNs) return false;
    }
}

// Debug timeline (-DZKSC_TAIL_TRACE, tools/trace_tail.py): thread 0 of CTA 0 and of CTA 1 of group 0 record %globaltimer at the
// phase boundaries of every round into g_tail_trace[round][cta][phase]; read back with zksc_debug_tail_trace.  Off by default.
#ifdef ZKSC_TAIL_TRACE
constexpr int kTraceRounds = 64, kTracePhases = 8;
__device__ unsigned long long g_tail_trace[kTraceRounds * 2 * kTracePhases];
#define ZKSC_TRACE(phase)                                                                                                              \
    do {                                                                                                                               \
        if (threadIdx.x == 0 && cta < 2 && group == 0 && round < (unsigned int)kTraceRounds)                                           \
            g_tail_trace[(round * 2 + cta) * kTracePhases + (phase)] = global_timer_ns();                                              \
    } while (0)
#else
#define ZKSC_TRACE(phase) do { } while (0)
#endif

// One CTA of the group of (proof, product), degree D: all rounds.
template <int D>
ZKSC_DEV void tail_body(const TailArgs& args, const int proof, const int product, FoldTab& s_tab, int& s_state) {
    constexpr int NL = Lazy<D>::NL;
    constexpr int NP = D;                       // point 1 is derived on the host (SKIP1)
    __shared__ Acc<NL> s_warp[kTailWarps][NP];
    __shared__ Fr s_tot[NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int cta = blockIdx.x, n_ctas = gridDim.x;
    const unsigned int group = proof * args.n_products + product;
    const unsigned int koff = args.koff[product];
    const Fr* in = args.in + (size_t)proof * args.in_proof_stride + (size_t)koff * args.in_tab_stride;
    unsigned long long in_stride = args.in_tab_stride;
    Fr* const out = args.out + (size_t)proof * args.out_proof_stride + (size_t)koff * args.out_tab_stride;
    const unsigned long long* mail = args.mail + (size_t)proof * kMailUnits;
    unsigned long long* relay = args.relay + (size_t)group * kMailUnits;
    unsigned long long* sums = args.sums + (size_t)group * n_ctas * kTailMaxDegree * 8;
    const unsigned int elem0 = proof * args.n_evals + args.eoff[product];      // first evaluation of this product
    unsigned long long* res = args.results + (size_t)elem0 * 8;
    unsigned long long half = args.half;

    for (unsigned int round = 0; round < args.n_rounds; round++, half >>= 1) {
        const unsigned int seq = args.seq0 + round;
        unsigned long long want = (half + kTailThreads - 1) / kTailThreads;
        const unsigned int n_active = want < n_ctas ? (unsigned int)want : n_ctas;
        if (cta >= n_active) return;            // no share of the table in this or any later round
        ZKSC_TRACE(0);
        // ---- 1. the challenge of the previous round
        if (warp == 0) {
            // only CTA 0 decides that the host has gone away (and then tells the others through the relay)
            const int st = tail_wait_mail(cta == 0 ? mail : relay, seq, s_tab, cta == 0 ? kTailTimeoutNs : 4 * kTailTimeoutNs);
            if (cta == 0 && n_active > 1) {
                const uint32_t* w = &s_tab.w[0][0];
                __syncwarp();
                const unsigned int tag = (st == 0) ? seq : kTailAbort;
                st_unit(relay + 2 * lane, w[2 * lane], tag);
                st_unit(relay + 2 * lane + 1, w[2 * lane + 1], tag);
            }
            if (lane == 0) s_state = st;
        }
        __syncthreads();
        if (s_state != 0) {
            if (s_state == 2 && cta == 0 && threadIdx.x < 8 * (D + 1)) st_unit(res + threadIdx.x, 0u, kTailTimeout);
            return;
        }
        ZKSC_TRACE(1);
        if (n_ctas > 1) fence_acq_rel_gpu();     // acquire: tables written by other CTAs before they handed in their sums
        ZKSC_TRACE(2);
        // ---- 2. fold + evaluate this CTA's share
        Acc<NL> acc[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) acc_zero(acc[p]);
        for (unsigned long long x = (unsigned long long)cta * kTailThreads + threadIdx.x; x < half; x += (unsigned long long)n_active * kTailThreads) {
            Fr a[D], b[D];
#pragma unroll
            for (int k = 0; k < D; k++) {
                const Fr* t = in + (size_t)k * in_stride;
                const Fr p0 = ld256_cg(t + x), p1 = ld256_cg(t + x + 2 * half);
                const Fr q0 = ld256_cg(t + x + half), q1 = ld256_cg(t + x + 3 * half);
                a[k] = fr_fold_tab(p0, p1, s_tab);
                b[k] = fr_fold_tab(q0, q1, s_tab);
                Fr* o = out + (size_t)k * args.out_tab_stride;
                st256(o + x, a[k]);
                st256(o + x + half, b[k]);
            }
            accumulate_points<D, true>(acc, a, b, D + 1);
        }
        ZKSC_TRACE(3);
        if (n_active > 1) fence_acq_rel_gpu();   // release: this thread's table stores, before the CTA's sums go out
        ZKSC_TRACE(4);
        // ---- 3. reduce inside the CTA
#pragma unroll
        for (int p = 0; p < NP; p++) {
            acc_warp_reduce(acc[p]);
            if (lane == 0) s_warp[warp][p] = acc[p];
        }
        __syncthreads();      // (s_tab may be rewritten from here on)
        bool ok = true;
        static_assert(NP <= kTailWarps, "one warp per evaluation point");
        if (warp < NP) {
            // warp p finishes point p: the NP reductions (each ends in two Montgomery products) run side by side instead of one
            // after the other in warp 0 -- this step was 3.4-3.8 us of every resident round (profiles/r01_tail_trace_c2_v10.txt)
            const int p = warp;
            Acc<NL> a;
            if (lane < kTailWarps) a = s_warp[lane][p];
            else acc_zero(a);
            acc_warp_reduce(a);
            const Fr v = acc_finish<NL>(a);                    // lane 0 holds the CTA's total
            if (n_active > 1) {
#pragma unroll
                for (int l = 0; l < 8; l++) {
                    const uint32_t limb = __shfl_sync(0xffffffffu, v.l[l], 0);
                    if (lane == l) st_unit(sums + ((size_t)cta * kTailMaxDegree + p) * 8 + l, limb, seq);
                }
            } else if (lane == 0) {
                s_tot[p] = v;
            }
        }
        in = out;
        in_stride = args.out_tab_stride;
        ZKSC_TRACE(5);
        if (cta != 0) continue;
        // ---- 4. CTA 0: the other CTAs' sums (warp p collects point slot p)
        if (n_active > 1) {
            for (int p = warp; p < NP; p += kTailWarps) {
                Acc<9> a;
                acc_zero(a);
                for (unsigned int c = lane; c < n_active; c += 32) {
                    Fr v;
                    ok = tail_read_elem(sums + ((size_t)c * kTailMaxDegree + p) * 8, seq, v) && ok;
                    acc_add<9, 8>(a, v.l);
                }
                acc_warp_reduce(a);
                if (lane == 0) s_tot[p] = acc9_reduce(a);
            }
            fence_acq_rel_gpu();               // acquire the other CTAs' table stores; the relay / host messages release them
        }
        __syncthreads();
        ZKSC_TRACE(6);
        // ---- 5. sharded contexts: all-to-all of the group's sums through peer memory, modular sum
        if (args.n_ranks > 1) {
            const unsigned int G = args.n_ranks, slot = (seq & 1u) * G;
            if (warp == 0) {
                for (int p = 0; p < NP; p++) {
                    const unsigned int elem = elem0 + (p == 0 ? 0 : p + 1);
                    const uint32_t limb = s_tot[p].l[lane & 7];
                    for (unsigned int g = lane >> 3; g < G; g += 4)
                        st_unit(args.peer_units[g] + ((size_t)(slot + args.rank) * args.xch_cap + elem) * 8 + (lane & 7), limb, seq);
                }
            }
            __syncthreads();
            for (int p = warp; p < NP; p += kTailWarps) {
                const unsigned int elem = elem0 + (p == 0 ? 0 : p + 1);
                Acc<9> a;
                acc_zero(a);
                if (lane < (int)G) {
                    Fr v;
                    ok = tail_read_elem(args.peer_units[args.rank] + ((size_t)(slot + lane) * args.xch_cap + elem) * 8, seq, v) && ok;
                    acc_add<9, 8>(a, v.l);
                }
                acc_warp_reduce(a);
                if (lane == 0) s_tot[p] = acc9_reduce(a);
            }
            __syncthreads();
        }
        // ---- 6. publish
        if (__syncthreads_or(!ok)) {
            if (threadIdx.x < 8 * (D + 1)) st_unit(res + threadIdx.x, 0u, kTailFailed);
            return;
        }
        if (threadIdx.x < 8 * NP) {
            const int p = threadIdx.x >> 3, l = threadIdx.x & 7;
            st_unit(res + (p == 0 ? 0 : p + 1) * 8 + l, s_tot[p].l[l], seq);
        }
        ZKSC_TRACE(7);
    }
}

__global__ void __launch_bounds__(kTailThreads, 1) tail_kernel(const __grid_constant__ TailArgs args) {
    __shared__ FoldTab s_tab;
    __shared__ int s_state;
    const int proof = blockIdx.y, product = blockIdx.z;
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
