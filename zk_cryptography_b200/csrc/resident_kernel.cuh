// K4 -- the resident rounds kernel: every round of a proof from the one where launch latency starts to matter to the last.
//
// An ordinary round costs a launch, a cold start, a cross-block reduction whose last block walks the per-block partials and a
// completion word: ~20 us on top of the arithmetic, every round (profiles/r02_bench_c2_v0_round1_kernels.json: 2^19 pairs
// 66 us against 31 us of HBM time).  This kernel is launched ONCE (cooperatively: all of its CTAs are guaranteed resident) with
// the CTA shape of the round kernels -- 128 threads, <= 128 registers, 4 CTAs per SM -- and loops over the remaining rounds:
//
//   the first CTA of a proof polls the previous round's challenge (its fold table, posted by the host into pinned memory) and
//     relays it through HBM to the other CTAs working on that proof (release / acquire on a sequence word);
//   every CTA folds its share of every table in place and accumulates the next round's evaluations -- the arithmetic of
//     round_kernel<D, FOLD = true, SKIP1 = true> (kernels.cuh), the fold table read from shared memory with LDS.128;
//   CTA-level reduction, per-CTA partial to HBM, arrival counter; the last CTA of a (proof, product) group to arrive sums the
//     partials (all 128 threads, loads issued together), and on a sharded context exchanges the group's sums with the other ranks
//     through NVLink peer memory;
//   it then publishes the evaluations into pinned host memory.  The host thread absorbs the round polynomial into the
//     Fiat-Shamir transcript (which never leaves the host: north_star), derives the challenge, posts its fold table.
//
// CTAs whose share of the (shrinking) table is empty leave.  No grid-wide barrier is needed: a CTA can only receive round j+1's
// challenge after EVERY CTA's round-j stores were fenced and counted, which orders the in-place table updates across CTAs.
//
// Sharded contexts (rank g holds the entries i = g mod G): when the table is down to `gather_total` entries over all ranks, every
// rank copies its folded shard into a staging area that its peers can read (CUDA IPC), and in the next round every rank PULLS
// the whole table over NVLink, folds it into local memory and carries on replicated -- the last log2(gather_total) rounds need
// no exchange at all, and there is no NCCL call, extra launch or second resident kernel for the residual.
//
// Host <-> device messages are self-validating 64-bit units  payload | sequence number << 32  (single 64-bit accesses): a reader
// that finds the expected sequence number in every unit has the whole message, whatever order PCIe or NVLink delivered it in.
//
// Replaces, per round, multi_composed_sumcheck.rs:81-89 + :103-105 of the reference, like the round kernels.
#pragma once
#include "kernels.cuh"

namespace zksc {

constexpr int kResThreads = kThreads;               // 128
constexpr int kResWarps = kResThreads / 32;
constexpr int kResMaxDegree = 5;                    // table fold (fr.cuh mul_fixed_rows) degrees
constexpr int kResMaxProducts = 8;                  // == ZKSC_MAX_PRODUCTS
constexpr int kMailUnits = 64;                      // fold-table words per proof and round
constexpr unsigned int kTailAbort = 0xffffffffu;    // sequence number that tells the kernel to leave
constexpr unsigned int kTailTimeout = 0xfffffffeu;  // published when no challenge arrived in time (nothing folded: recoverable)
constexpr unsigned int kTailFailed = 0xfffffffdu;   // published when a CTA or a peer GPU went missing in the middle of a round
constexpr unsigned long long kTailTimeoutNs = 1000ull * 1000 * 1000;   // a proof's first CTA waiting for the host; everything else waits a multiple
constexpr unsigned int kNoGather = 0xffffffffu;
// At most this many CTAs of a proof poll the host mailbox themselves.  Measured (profiles/r02_resident_v2_direct_poll.txt): k CTAs
// polling pinned host memory cost ~2 us of PCIe contention EACH per round and delay the publication of the round before, so it is
// one -- the proof's first CTA alone -- and everyone else listens to its relay in HBM.
constexpr unsigned int kResDirectPoll = 1;
#ifndef ZKSC_RES_MINB
#define ZKSC_RES_MINB 4
#endif

struct ResArgs {
    const Fr* in;              // tables before the first fold of this kernel (table 0 of proof 0)
    Fr* out;                   // folded tables (may alias `in`)
    unsigned long long in_tab_stride, in_proof_stride, out_tab_stride, out_proof_stride;   // elements
    unsigned long long half;   // pairs per (local) table of the first round evaluated here
    unsigned int n_rounds;     // rounds to run
    unsigned int seq0;         // sequence number of the first round's messages
    unsigned int n_proofs, n_products, n_evals, n_tables;   // n_evals = sum of (degree + 1), n_tables = sum of degrees, per proof
    unsigned int deg[kResMaxProducts], koff[kResMaxProducts], eoff[kResMaxProducts];
    unsigned int cpg;          // CTAs per (proof, product) group; group g owns CTAs [g cpg, (g + 1) cpg)
    const unsigned long long* mail;   // [proof][kMailUnits]       host-mapped: fold-table word (FoldTabS order) | seq << 32
    unsigned long long* results;      // [proof][n_evals][8]       host-mapped: limb | seq << 32
    unsigned long long* status;       // [proof]                   host-mapped: seq | kTailTimeout << 32 ("gave up waiting for round seq's challenge;
                                      //                           nothing of it was folded") or seq | kTailFailed << 32.  Kept apart from the results: a
                                      //                           round published BEFORE the time-out must stay readable (synchronous launches, e.g. under
                                      //                           Nsight Compute, let the host look only after the kernel has ended)
    unsigned long long* relay;        // [2][relay_cap][kMailUnits]  HBM: the fold table of round seq of proof b in slot [seq & 1][b], as units
    unsigned long long* relay_tags;   // [2][relay_cap]              HBM: seq | status << 32 once the table of round seq is complete (status 0) or the
                                      //                             proof's first CTA left instead (1 = told to, 2 = timed out); other seqs are stale
    unsigned int relay_cap;           // proofs the relay was allocated for (the layout must not depend on the handle in use)
    Fr* partials;                     // [group][kResMaxDegree][cpg]  HBM
    unsigned int* counters;           // [group]                   HBM, zero between rounds
    unsigned int* work;               // [group][warp in CTA][kWorkCtrWords]  HBM, zero between rounds: chunks of 32 pairs handed out beyond every
                                      //                           warp's first (nullptr: static assignment only)
    // cross-GPU exchange (n_ranks > 1): every rank's unit buffer [2][n_ranks][xch_cap][8]
    unsigned long long* peer_units[kMaxRanks];
    unsigned int n_ranks, rank, xch_cap;
    unsigned int host_reduce;         // n_ranks > 1 inside one process: no exchange, every rank publishes its own sums and the host adds them
    // gather (n_ranks > 1): after round `gather_round` (counted from this kernel's first) the folded local tables -- gather_local
    // entries each -- are copied to this rank's stage; round gather_round + 1 reads all ranks' stages and writes `tail`
    unsigned int gather_round;
    unsigned long long gather_local;
    const Fr* peer_stage[kMaxRanks];  // [proof][table][gather_local] of every rank (own entry: local address)
    Fr* tail;                         // [proof][table][gather_local * n_ranks / 2]  local, replicated rounds
    unsigned long long tail_tab_stride, tail_proof_stride;
};

#ifdef ZKSC_RES_IMPL   // the kernel itself: res_inst.cu only (zksc.cu needs just the declarations above)
ZKSC_DEV void ld_units2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
ZKSC_DEV void st_unit(unsigned long long* p, unsigned int payload, unsigned int seq) {
    const unsigned long long v = (unsigned long long)payload | ((unsigned long long)seq << 32);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
ZKSC_DEV void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }     // MEMBAR.ALL.GPU (__threadfence() is the sequentially consistent, slower one)
ZKSC_DEV void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
ZKSC_DEV unsigned int ld_cg_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
ZKSC_DEV void st_release_gpu(unsigned long long* p, unsigned long long v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
ZKSC_DEV unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// relay tag of round seq: status 0 = the table is there, 1 = abort, 2 = time-out
ZKSC_DEV unsigned long long relay_tag(unsigned int seq, unsigned int status) { return (unsigned long long)seq | ((unsigned long long)status << 32); }
// an element in peer-visible (CUDA IPC) memory: 128-bit accesses only (kernels.cuh st256_2x128), system-scope loads
ZKSC_DEV Fr ld256_sys_2x128(const Fr* p) {
    Fr v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]) : "l"(p) : "memory");
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7]) : "l"(p) : "memory");
    return v;
}

// Warp-level: poll a 64-unit mailbox (two units per lane) until every unit carries `seq`; the fold table lands in `tab`.
// Returns 0 = ok, 1 = abort requested (unit 0 carries kTailAbort), 2 = timed out.  `watch` (optional): a relay tag through which the
// proof's first CTA announces that IT gave up or was told to leave -- the only CTA whose time-out counts (see the host's tail_post).
ZKSC_DEV int res_wait_mail(const unsigned long long* mail, unsigned int seq, FoldTabS& tab, unsigned long long timeout_ns, const unsigned long long* watch = nullptr) {
    const int lane = threadIdx.x & 31;
    const unsigned long long t0 = global_timer_ns();
    for (unsigned int spins = 1;; spins++) {
        unsigned long long a, b;
        ld_units2(mail + 2 * lane, a, b);
        const bool ok = ((unsigned int)(a >> 32) == seq) && ((unsigned int)(b >> 32) == seq);
        const bool abort = (lane == 0) && ((unsigned int)(a >> 32) == kTailAbort);
        if (__any_sync(0xffffffffu, abort)) return 1;
        if (__all_sync(0xffffffffu, ok)) {
            tab.w[2 * lane] = (unsigned int)a;
            tab.w[2 * lane + 1] = (unsigned int)b;
            return 0;
        }
        if ((spins & 0x3ff) == 0) {
            int st = 0;
            if (lane == 0) {
                if (global_timer_ns() - t0 > timeout_ns) st = 2;
                if (watch) {
                    const unsigned long long w = ld_acquire_gpu(watch);
                    if ((unsigned int)w == seq && (w >> 32) != 0) st = (int)(w >> 32);
                }
            }
            st = __shfl_sync(0xffffffffu, st, 0);
            if (st) return st;
        }
    }
}
// One lane: read an element published as 8 units; spins until all carry `seq`.  false on timeout.
ZKSC_DEV bool res_read_elem(const unsigned long long* u, unsigned int seq, Fr& out) {
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

// Debug timeline (-DZKSC_RES_TRACE, tools/trace_resident.py): %globaltimer at the phase boundaries of every round, recorded by thread 0 of
// group 0's first CTA (phases 0-3) and of whichever CTA arrives last (phases 4-7); read back with zksc_debug_res_trace.  Off by default.
#ifdef ZKSC_RES_TRACE
constexpr int kTraceRounds = 64, kTracePhases = 8;
__device__ unsigned long long g_res_trace[kTraceRounds * kTracePhases];
// ... and per CTA, for the kernel's first four rounds: [round][cta] = {start (challenge in shared memory), all warps done with fold + evaluate, %smid}
constexpr int kTraceCtas = 640, kTraceCtaRounds = 4;
__device__ unsigned long long g_res_cta_trace[kTraceCtaRounds * kTraceCtas * 3];
#define ZKSC_TRACE_CTA(which)                                                                                             \
    do {                                                                                                                 \
        if (threadIdx.x == 0 && round < (unsigned int)kTraceCtaRounds && blockIdx.x < (unsigned int)kTraceCtas) {        \
            unsigned int smid_;                                                                                          \
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));                                                         \
            g_res_cta_trace[(round * kTraceCtas + blockIdx.x) * 3 + (which)] = global_timer_ns();                       \
            g_res_cta_trace[(round * kTraceCtas + blockIdx.x) * 3 + 2] = smid_;                                        \
        }                                                                                                                \
    } while (0)
#define ZKSC_TRACE(cond, phase)                                                                                          \
    do {                                                                                                                 \
        if (threadIdx.x == 0 && (cond) && group == 0 && round < (unsigned int)kTraceRounds)                          \
            g_res_trace[round * kTracePhases + (phase)] = global_timer_ns();                                             \
    } while (0)
#else
#define ZKSC_TRACE(cond, phase) do { } while (0)
#define ZKSC_TRACE_CTA(which) do { } while (0)
#endif

// fold + evaluate one pair of every factor; the folded entries go to o[k] + x and o[k] + x + half
template <int D, class LOAD, class A, int NP>
ZKSC_DEV void res_pair(A (&acc)[NP], LOAD&& load, Fr* out, unsigned int out_stride, unsigned int x, unsigned int half, const FoldTabS& tab) {
#ifndef ZKSC_RES_SEMI3
#define ZKSC_RES_SEMI3 0
#endif
    constexpr bool kSemi = (D == 2) || (ZKSC_RES_SEMI3 && D == 3);     // see fr_fold_tab
    Fr a[D], b[D];
    if constexpr (D == 2) {
        // all eight entries of the pair in flight at once: one exposed memory latency per iteration instead of two
        const Fr p0 = load(0, x), p1 = load(0, x + 2 * half), q0 = load(0, x + half), q1 = load(0, x + 3 * half);
        const Fr r0 = load(1, x), r1 = load(1, x + 2 * half), s0 = load(1, x + half), s1 = load(1, x + 3 * half);
        a[0] = fr_fold_tab<kSemi>(p0, p1, tab);
        b[0] = fr_fold_tab<kSemi>(q0, q1, tab);
        st256(out + x, a[0]);
        st256(out + x + half, b[0]);
        a[1] = fr_fold_tab<kSemi>(r0, r1, tab);
        b[1] = fr_fold_tab<kSemi>(s0, s1, tab);
        st256(out + out_stride + x, a[1]);
        st256(out + out_stride + x + half, b[1]);
    } else {
#pragma unroll
        for (int k = 0; k < D; k++) {
            // T_{j-1} has 4 * half entries; its pairs are (y, y + 2 * half)
            const Fr p0 = load(k, x), p1 = load(k, x + 2 * half);
            const Fr q0 = load(k, x + half), q1 = load(k, x + 3 * half);
            a[k] = fr_fold_tab<kSemi>(p0, p1, tab);
            b[k] = fr_fold_tab<kSemi>(q0, q1, tab);
            Fr* o = out + (size_t)k * out_stride;
            st256(o + x, a[k]);
            st256(o + x + half, b[k]);
        }
    }
    accumulate_points<D, true>(acc, a, b, D + 1);
}

// All remaining rounds of one CTA of the group (proof, product), degree D.
// What identifies this CTA's work -- needed between the hot loops only, so it lives in shared memory and is re-read where it is
// used: the registers it would occupy across the fold + evaluate loop are what lets ptxas keep all loads of a pair in flight.
struct ResIds {
    unsigned int proof, product, ci, group, koff, elem0, poller;
};
template <int D>
ZKSC_DEV void res_rounds(const ResArgs& args, const volatile ResIds& ids, FoldTabS& s_tab, int& s_state, int& s_last) {
    constexpr int NL = Lazy<D>::NL;
    constexpr int NP = D;                        // point 1 is derived on the host (SKIP1)
    constexpr bool kSmemAcc = (D == 3);          // kernels.cuh: the d = 3 fold keeps its sums in shared memory (registers)
    __shared__ Acc<NL> s_warp[kResWarps][NP];
    __shared__ Acc<9> s_red[NP][kResWarps];
    __shared__ Fr s_tot[NP];
    __shared__ uint32_t s_acc[kSmemAcc ? NP * NL * kResThreads : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#define proof (ids.proof)
#define product (ids.product)
#define ci (ids.ci)
#define group (ids.group)
#define koff (ids.koff)
#define elem0 (ids.elem0)
#define poller (ids.poller != 0u)
    // 32-bit indices throughout (the host starts this kernel only for tables below 2^31 entries): fewer live registers in the hot loop
    const Fr* in = args.in + (size_t)proof * args.in_proof_stride + (size_t)koff * args.in_tab_stride;
    unsigned int in_stride = (unsigned int)args.in_tab_stride;
    Fr* out = args.out + (size_t)proof * args.out_proof_stride + (size_t)koff * args.out_tab_stride;
    unsigned int out_stride = (unsigned int)args.out_tab_stride;
    unsigned int half = (unsigned int)args.half;

    for (unsigned int round = 0; round < args.n_rounds; round++, half >>= 1) {
        const unsigned int seq = args.seq0 + round;
        const bool pull = (args.n_ranks > 1 && round == args.gather_round + 1);          // this round starts from every rank's stage
        if (pull) half = (unsigned int)(args.gather_local * args.n_ranks / 4);           // ... a table of gather_local * n_ranks entries
        const bool exchange = (args.n_ranks > 1 && !args.host_reduce && (args.gather_round == kNoGather || round <= args.gather_round));
        const unsigned int want = (half + kResThreads - 1) / kResThreads;
        const unsigned int n_active = want < args.cpg ? want : args.cpg;
        if (ci >= n_active) {
            // No share of the table in this round.  Tables only shrink -- except once on a sharded context, when the gathered table
            // is G times a shard: the CTAs that round will need skip ahead to it (and wait there for its challenge), the others leave.
            if (args.n_ranks > 1 && args.gather_round != kNoGather && round <= args.gather_round) {
                const unsigned int pull_want = ((unsigned int)(args.gather_local * args.n_ranks / 4) + kResThreads - 1) / kResThreads;
                if (ci < (pull_want < args.cpg ? pull_want : args.cpg)) {
                    round = args.gather_round;      // the loop's increment makes it the pull round
                    continue;
                }
            }
            return;
        }
        // ---- 1. the challenge of the previous round
        {
            // Few CTAs at work on this proof: each polls the host's mailbox itself (one PCIe read per poll and CTA).  Many: the
            // proof's first CTA polls and relays the table through HBM as self-validating units behind a sequence word -- the
            // word says when to look, the units say whether the table is complete, so the relay costs no fence.
            const bool direct = (n_active * args.n_products <= kResDirectPoll);
            unsigned long long* runits = args.relay + ((size_t)(seq & 1u) * args.relay_cap + proof) * kMailUnits;
            unsigned long long* rtag = args.relay_tags + (size_t)(seq & 1u) * args.relay_cap + proof;
            if (warp == 0) {
                int st;
                if (poller) {
                    st = res_wait_mail(args.mail + (size_t)proof * kMailUnits, seq, s_tab, kTailTimeoutNs);
                    if (!direct && st == 0) {
                        st_unit(runits + 2 * lane, s_tab.w[2 * lane], seq);
                        st_unit(runits + 2 * lane + 1, s_tab.w[2 * lane + 1], seq);
                        __syncwarp();
                    }
                    if (lane == 0 && (!direct || st != 0)) st_release_gpu(rtag, relay_tag(seq, (unsigned int)st));
                } else if (direct) {
                    st = res_wait_mail(args.mail + (size_t)proof * kMailUnits, seq, s_tab, 8 * kTailTimeoutNs, rtag);
                    if (st == 2 && ld_acquire_gpu(rtag) != relay_tag(seq, 2u)) st = 3;     // the first CTA did not time out, yet no challenge came
                } else {
                    // the relayed units validate themselves, so the listeners poll them directly (one L2 round trip less than waiting
                    // for the tag first); the tag is looked at now and then, for the first CTA's notice that it left
                    st = res_wait_mail(runits, seq, s_tab, 8 * kTailTimeoutNs, rtag);
                    if (st == 2 && ld_acquire_gpu(rtag) != relay_tag(seq, 2u)) st = 3;
                }
                if (lane == 0) s_state = st;
            }
        }
        __syncthreads();
        ZKSC_TRACE(ci == 0, 0);        // challenge in shared memory
        ZKSC_TRACE_CTA(0);
        if (s_state != 0) {
            // a timeout is the proof's first CTA's decision alone (the other CTAs only ever hear it through the relay): nothing of
            // this round has been folded anywhere when it is published
            if (s_state == 2 && poller && threadIdx.x == 0) st_unit(args.status + proof, seq, kTailTimeout);
            if (s_state == 3 && ci == 0 && threadIdx.x == 0) st_unit(args.status + proof, seq, kTailFailed);
            return;
        }
        const unsigned int x0 = ci * kResThreads + threadIdx.x, xs = n_active * kResThreads;
        if (pull) {
            // Entry i of the gathered table lives on rank i mod G at local index i / G of that rank's stage (128-bit system-scope
            // loads: peer memory over NVLink).  Every thread fetches exactly the entries it is about to fold into the local
            // `tail` buffer, which the rounds from here on update in place.
            const unsigned int G = args.n_ranks;
            const size_t tab0 = ((size_t)proof * args.n_tables + koff) * args.gather_local;
            in = out = args.tail + (size_t)proof * args.tail_proof_stride + (size_t)koff * args.tail_tab_stride;
            in_stride = out_stride = (unsigned int)args.tail_tab_stride;
            for (unsigned int x = x0; x < half; x += xs)
#pragma unroll 1
                for (int k = 0; k < D; k++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const unsigned int i = x + q * half;
                        st256(out + (size_t)k * out_stride + i, ld256_sys_2x128(args.peer_stage[i % G] + tab0 + (size_t)k * args.gather_local + i / G));
                    }
            __syncthreads();
        }
        // ---- 2. fold + evaluate this CTA's share
        Acc<NL> acc[NP];
        SmemAcc<NL> sacc[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            acc_zero(acc[p]);
            if constexpr (kSmemAcc) {
                sacc[p].p = s_acc + (size_t)p * NL * kResThreads + threadIdx.x;
#pragma unroll
                for (int i = 0; i < NL; i++) sacc[p].p[i * kResThreads] = 0u;
            }
        }
        {
            const Fr* tin = in;
            const unsigned int tstride = in_stride;
            auto load = [&](int k, unsigned int i) { return ld256_cg(tin + (size_t)k * tstride + i); };
            // More than one pair per thread: the warps of the group take chunks of 32 pairs from a shared counter instead of a fixed
            // stride.  The CTAs of an SM do not progress at the same pace (the warp schedulers favour the older ones: with a fixed
            // split the first CTA of an SM was done after 64 us of an 85 us round, the fourth after 81 us, and the SM ran half empty
            // in between -- profiles/r02_resident_cta_spread.txt), nor do all SMs; handing the work out keeps every warp busy to
            // the end.  The next chunk is asked for before the current one is worked on, so the atomic's latency is hidden.  The
            // rounds that copy entries for other ranks or from them keep the fixed split (a thread must fold what it copied).
            // One counter per warp index in the CTA (warp w works on the chunks c = w mod 4), each in its own cache line: atomics on a
            // single address go through at ~0.8 per ns.  Degree 1 is bound by HBM alone: nothing to balance.
            const bool dynamic = D >= 2 && args.work != nullptr && want > args.cpg && !pull && !(args.n_ranks > 1 && round == args.gather_round);
            if (dynamic) {
                const unsigned int n_chunks = (half + 31u) >> 5;
                unsigned int* ctr = args.work + ((size_t)group * kResWarps + warp) * kWorkCtrWords;
                for (unsigned int i = ci;;) {
                    const unsigned int c = i * kResWarps + warp;
                    if (c >= n_chunks) break;
                    unsigned int nxt = 0;
                    if (lane == 0) nxt = atomicAdd(ctr, 1u) + n_active;
                    const unsigned int x = (c << 5) + lane;
                    if (x < half) {
                        if constexpr (kSmemAcc) res_pair<D>(sacc, load, out, out_stride, x, half, s_tab);
                        else res_pair<D>(acc, load, out, out_stride, x, half, s_tab);
                    }
                    i = __shfl_sync(0xffffffffu, nxt, 0);
                }
            } else {
                for (unsigned int x = x0; x < half; x += xs) {
                    if constexpr (kSmemAcc) res_pair<D>(sacc, load, out, out_stride, x, half, s_tab);
                    else res_pair<D>(acc, load, out, out_stride, x, half, s_tab);
                }
            }
        }
        if constexpr (kSmemAcc) {
#pragma unroll
            for (int p = 0; p < NP; p++)
#pragma unroll
                for (int i = 0; i < NL; i++) acc[p].l[i] = sacc[p].p[i * kResThreads];
        }
        ZKSC_TRACE(ci == 0, 1);        // fold + evaluate done
#ifdef ZKSC_RES_TRACE
        __syncthreads();
        ZKSC_TRACE_CTA(1);
#endif
        if (args.n_ranks > 1 && round == args.gather_round) {
            // the folded shard, where the peers can read it (the next round pulls from every rank's stage)
            Fr* stage = const_cast<Fr*>(args.peer_stage[args.rank]) + ((size_t)proof * args.n_tables + koff) * args.gather_local;
            for (unsigned int x = x0; x < half; x += xs)
#pragma unroll 1
                for (int k = 0; k < D; k++) {
                    st256_2x128(stage + (size_t)k * args.gather_local + x, ld256(out + (size_t)k * out_stride + x));               // this thread's own stores
                    st256_2x128(stage + (size_t)k * args.gather_local + x + half, ld256(out + (size_t)k * out_stride + x + half));
                }
            fence_acq_rel_sys();
        }
        in = out;
        in_stride = out_stride;
        // ---- 3. reduce inside the CTA
#pragma unroll
        for (int p = 0; p < NP; p++) {
            acc_warp_reduce(acc[p]);
            if (lane == 0) s_warp[warp][p] = acc[p];
        }
        __syncthreads();      // (s_tab may be rewritten from here on)
        for (int p = warp; p < NP; p += kResWarps) {
            Acc<NL> a;
            if (lane < kResWarps) a = s_warp[lane][p];
            else acc_zero(a);
            acc_warp_reduce(a);
            const Fr v = acc_finish_warp<NL>(a);
            if (lane == 0) {
                if (n_active > 1) st256(args.partials + ((size_t)group * kResMaxDegree + p) * args.cpg + ci, v);
                else s_tot[p] = v;
            }
        }
        __syncthreads();
        ZKSC_TRACE(ci == 0, 2);        // CTA-level reduction done
        if (n_active > 1) {
            if (threadIdx.x == 0) {
                fence_acq_rel_gpu();          // this CTA's table stores and its partial, before it is counted
                s_last = (atomicAdd(args.counters + group, 1u) == n_active - 1);
            }
            __syncthreads();
            ZKSC_TRACE(ci == 0, 3);    // counted
            if (!s_last) continue;
            ZKSC_TRACE(true, 4);           // last to arrive
            // ---- 4. the last CTA of the group to arrive: sum the partials (all threads, loads issued together)
            fence_acq_rel_gpu();
            cta_sum_points<NP>(args.partials + (size_t)group * kResMaxDegree * args.cpg, args.cpg, 1, n_active, s_red);
            __syncthreads();
            for (int p = warp; p < NP; p += kResWarps) {
                const Fr v = warp_finish_sum<NP>(s_red[p]);
                if (lane == 0) s_tot[p] = v;
            }
            if (threadIdx.x == 0) {
                args.counters[group] = 0u;     // nobody arrives for the next round before its challenge exists
            }
            if (args.work && threadIdx.x < kResWarps) args.work[((size_t)group * kResWarps + threadIdx.x) * kWorkCtrWords] = 0u;
            __syncthreads();
            ZKSC_TRACE(true, 5);           // partials summed
        }
        bool ok = true;
        // ---- 5. sharded contexts: all-to-all of the group's sums through peer memory, modular sum
        if (exchange) {
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
            for (int p = warp; p < NP; p += kResWarps) {
                const unsigned int elem = elem0 + (p == 0 ? 0 : p + 1);
                Fr v = fr_zero();
                if (lane < (int)G) ok = res_read_elem(args.peer_units[args.rank] + ((size_t)(slot + lane) * args.xch_cap + elem) * 8, seq, v) && ok;
                // G <= 8 canonical values: a shuffle tree of modular additions (a Montgomery reduction of the plain sum costs ~1 us of
                // dependent instructions in a lone warp)
#pragma unroll
                for (int d = kMaxRanks / 2; d >= 1; d >>= 1) {
                    Fr o;
#pragma unroll
                    for (int i = 0; i < 8; i++) o.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
                    v = fr_add(v, o);
                }
                if (lane == 0) s_tot[p] = v;
            }
            __syncthreads();
        }
        // ---- 6. publish
        unsigned long long* res = args.results + (size_t)elem0 * 8;
        if (__syncthreads_or(!ok)) {
            if (threadIdx.x == 0) st_unit(args.status + proof, seq, kTailFailed);
            return;
        }
        if (threadIdx.x < 8 * NP) {
            const int p = threadIdx.x >> 3, l = threadIdx.x & 7;
            st_unit(res + (p == 0 ? 0 : p + 1) * 8 + l, s_tot[p].l[l], seq);
        }
        ZKSC_TRACE(true, 6);               // published
    }
#undef proof
#undef product
#undef ci
#undef group
#undef koff
#undef elem0
#undef poller
}

// DSEL > 0: every product has degree DSEL (registers are allocated for that degree alone); DSEL == 0: mixed degrees.
// Degrees 1 and 2 fit 128 registers (4 CTAs per SM); degree 3 spills there (208 bytes of stack, ~45 local loads and stores per pair)
// and runs 3 % faster on large rounds and 4.5 us faster on small ones with 168 registers and 3 CTAs per SM
// (profiles/r02_resident_d3_variants.txt); degrees 4, 5 and the mixed form likewise.
template <int DSEL>
__global__ void __launch_bounds__(kResThreads, (DSEL == 0 || DSEL >= 3) ? 3 : ZKSC_RES_MINB) resident_kernel(const __grid_constant__ ResArgs args) {
    __shared__ __align__(16) FoldTabS s_tab;
    __shared__ int s_state, s_last;
    __shared__ ResIds s_ids;
    {
        const unsigned int g = blockIdx.x / args.cpg, c = blockIdx.x % args.cpg;
        const unsigned int pr = g / args.n_products, pd = g % args.n_products;
        if (pr >= args.n_proofs) return;
        if (threadIdx.x == 0) {
            s_ids.proof = pr; s_ids.product = pd; s_ids.ci = c; s_ids.group = g;
            s_ids.koff = args.koff[pd];
            s_ids.elem0 = pr * args.n_evals + args.eoff[pd];
            s_ids.poller = (c == 0 && pd == 0) ? 1u : 0u;
        }
        __syncthreads();
    }
    const volatile ResIds& ids = s_ids;
    if constexpr (DSEL > 0) {
        res_rounds<DSEL>(args, ids, s_tab, s_state, s_last);
    } else {
        switch (args.deg[ids.product]) {
            case 1: res_rounds<1>(args, ids, s_tab, s_state, s_last); break;
            case 2: res_rounds<2>(args, ids, s_tab, s_state, s_last); break;
            case 3: res_rounds<3>(args, ids, s_tab, s_state, s_last); break;
            case 4: res_rounds<4>(args, ids, s_tab, s_state, s_last); break;
            case 5: res_rounds<5>(args, ids, s_tab, s_state, s_last); break;
            default: break;
        }
    }
}
#endif  // ZKSC_RES_IMPL

}  // namespace zksc
