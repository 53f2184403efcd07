// zksc engine: C ABI of include/zksc.h over the sm_100a kernels of kernels.cuh.
//
// Data layout in HBM: the reference's own element layout (32-byte Montgomery elements, array of
// elements per table) -- one LDG.E.256 / STG.E.256 per element per thread, 1 KiB contiguous per warp.
//   orig : [proof][table][N_local]     the caller's tables, never modified
//   work : [proof][table][N_local/2]   round >= 1 tables, folded in place from then on
// Round j >= 1 reads T_{j-1} once and writes T_j once (fold fused with the next round's evaluation).
#include "../../include/zksc.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "host_field.hpp"
#include "host_pairing.hpp"
#include "aux_kernels.cuh"
#include "kernels.cuh"
#include "resident_kernel.cuh"
#include "g1.cuh"

#if __has_include(<nccl.h>)
#include <nccl.h>
#define ZKSC_HAVE_NCCL_H 1
#else
#define ZKSC_HAVE_NCCL_H 0
#endif

using namespace zksc;
using zksc::host::FrH;

static_assert(kMaxDegree == ZKSC_MAX_DEGREE, "header / kernel limits differ");
static_assert(kResMaxProducts == ZKSC_MAX_PRODUCTS, "header / resident kernel limits differ");
static_assert(kWorkCtrBase > kMaxBatch * ZKSC_MAX_PRODUCTS, "work counters overlap the arrival counters");
cudaError_t zksc_launch_resident(int dsel, unsigned int ctas, cudaStream_t s, const ResArgs& a);   // res_inst.cu
int zksc_resident_occ(int dsel);

static thread_local std::string g_create_error;

constexpr size_t kXchTagBytes = 256;               // 2 * kMaxRanks tags, padded
constexpr unsigned int kXchCap = 2048;             // elements per (slot, rank): rounds with more partials use NCCL
constexpr size_t kXchStageElems = 1u << 15;        // gather stage of a rank: its folded shard of every table, read by the peers (1 MiB)
constexpr unsigned long long kGatherDefault = 2048;   // sharded contexts: gather the shards when a table is down to this many entries in total
// The resident kernel takes over from the round with at most this many limb products per rank (ZKSC_TAIL_WORK overrides): degree 2
// from 2^21 pairs, degree 3 from 2^20 (above that the ordinary launch's constant-bank fold table is worth more than its ~20 us of
// fixed cost).  Since the resident kernel hands its chunks out dynamically the cross-over sits one round earlier than with a fixed
// split (profiles/r02_resident_dynamic_ab.txt), at the same place as on sharded contexts, where an ordinary round also pays the
// cross-rank exchange in its last block.
constexpr unsigned long long kTailWorkDefault = 1100ull * 1000 * 1000;
constexpr unsigned long long kTailWorkSharded = 1100ull * 1000 * 1000;
// exchange buffer of a rank: tags[2][G] | data[2][G][kXchCap] elements (round kernels) | units[2][G][kXchCap][8] (resident kernel) |
// stage[kXchStageElems] elements (resident kernel, gather)
static inline size_t kXchUnitsOffset(int G) { return kXchTagBytes + (size_t)2 * G * kXchCap * 32; }
static inline size_t kXchUnitsBytes(int G) { return (size_t)2 * G * kXchCap * 64; }
static inline size_t kXchStageOffset(int G) { return kXchUnitsOffset(G) + kXchUnitsBytes(G); }

// ------------------------------------------------------------------------------------------------
// NCCL, loaded at run time (single-GPU use must not depend on libnccl being present)
// ------------------------------------------------------------------------------------------------
#if ZKSC_HAVE_NCCL_H
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GetErrorString) { err = "libnccl lacks required symbols"; return false; }
        return true;
    }
};
static NcclApi g_nccl;
#endif

// ------------------------------------------------------------------------------------------------
// Host worker pool for the per-proof half of a round when many independent proofs share a launch (c5: 64 transcripts per
// round).  Each proof's step -- interpolate, serialise, SHA-256, challenge, fold table, claim -- touches only that proof's
// state, so the proofs of a round are spread over a few threads.  Workers spin only while a zksc_prove call is in
// progress (a round's host step is ~100 us and must not pay a wake-up) and sleep on a condition variable otherwise.
class HostPool {
   public:
    explicit HostPool(int workers) {
        for (int i = 0; i < workers; i++) th_.emplace_back([this] { worker(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m_); quit_ = true; active_.store(false); }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void begin() { { std::lock_guard<std::mutex> lk(m_); active_.store(true, std::memory_order_release); } cv_.notify_all(); }
    void end() { active_.store(false, std::memory_order_release); }
    // fn(i) for i in [0, n); returns when all are done.  Only between begin() and end().
    void parallel_for(uint32_t n, const std::function<void(uint32_t)>& fn) {
        fn_ = &fn; n_ = n;
        next_.store(0, std::memory_order_relaxed);
        pending_.store((int)th_.size(), std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        run();
        while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
    }
    size_t workers() const { return th_.size(); }

   private:
    static void cpu_relax() {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    void run() {
        for (;;) {
            const uint32_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_) break;
            (*fn_)(i);
        }
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return quit_ || active_.load(std::memory_order_acquire); });
                if (quit_) return;
            }
            while (active_.load(std::memory_order_acquire)) {
                const uint64_t g = gen_.load(std::memory_order_acquire);
                if (g != seen) {
                    seen = g;
                    run();
                    pending_.fetch_sub(1, std::memory_order_acq_rel);
                } else {
                    cpu_relax();
                }
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    bool quit_ = false;
    std::atomic<bool> active_{false};
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0};
    std::atomic<uint32_t> next_{0};
    const std::function<void(uint32_t)>* fn_ = nullptr;
    uint32_t n_ = 0;
};
constexpr uint32_t kPoolMinProofs = 16;   // fewer proofs per round than this: the calling thread does them itself

struct zksc_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    uint64_t* gkr_stage = nullptr;        // pinned staging buffer of zksc_gkr_prove (one host->device copy per layer)
    size_t gkr_stage_cap = 0;             // in uint64_t
    HostPool* pool = nullptr;             // created on first use by a batched zksc_prove (ZKSC_HOST_THREADS, default up to 8 threads)
    bool pool_active = false;
    cudaStream_t copy_stream = nullptr;   // zksc_tables_reupload_begin: host->device copies that overlap the compute stream
    cudaEvent_t copy_event = nullptr;
    Fr* partials = nullptr;
    size_t partials_cap = 0;
    unsigned int* counters = nullptr;
    Fr* results_dev = nullptr;   // [n_ranks][kMaxBatch * kMaxEvals] (allgather target)
    Fr* results_send = nullptr;  // [kMaxBatch * kMaxEvals]
    Fr* results_host = nullptr;  // pinned, device-mapped: the final block of a round writes the evaluations straight into it
    Fr* results_host_dev = nullptr;   // device address of results_host
    volatile unsigned int* flag_host = nullptr;   // pinned, device-mapped completion word of the current round (system-scope store by the kernel)
    unsigned int* flag_dev = nullptr;
    unsigned int flag_seq = 0;
    bool mapped_results = true;  // ZKSC_NO_MAPPED=1: cudaMemcpyAsync + stream synchronize per round instead
    size_t results_cap = 0;      // elements per rank slot
    int occ[kMaxDegree + 1][6];   // resident CTAs per SM of round_kernel<D, variant> (0..2) and round_tma_kernel (3..5; 0 = none)
    bool staged = true;           // use the TMA-staged kernels where they apply (ZKSC_NO_STAGED=1 turns them off)
    bool staged_fold = false;     // ... also for the fused fold+evaluate rounds (ZKSC_STAGED_FOLD=1; slower today: DESIGN.md)
    std::string err;
    int rank = 0, n_ranks = 1;
    // Single-process multi-GPU (zksc_ctx_create_multi): the context the caller holds is a PARENT with one child context per device.
    // A child is an ordinary sharded context (rank g of G) whose per-round partial evaluations are not exchanged on the device but
    // published to its own pinned host buffer: the parent's host thread adds the G partials and runs the ONE transcript
    // (host_reduce).  The children's rounds run concurrently, one pool thread per device.
    std::vector<zksc_ctx*> kids;
    zksc_ctx* parent = nullptr;
    bool host_reduce = false;
    bool gpu_pool_active = false;
    HostPool* gpu_pool = nullptr;            // parent: kids.size() - 1 workers, spinning while a call is in progress
    unsigned long long gather_entries = 1;   // sharded contexts: the shards are gathered when a table is down to this many entries IN TOTAL
    // resident rounds kernel (resident_kernel.cuh): mailbox + result units in pinned, device-mapped host memory
    bool fuse_products = true;           // ZKSC_NO_FUSE=1: one launch per product even when the degrees agree
    bool tail_enabled = true;            // ZKSC_NO_TAIL=1: every round is its own launch
    unsigned int dyn_max_groups = 8;     // ZKSC_DYN_MAX_GROUPS (<= kDynMaxGroups): most groups of a launch that takes its chunks from counters
    bool round_dynamic = true;           // ZKSC_ROUND_STATIC=1: the round kernels split a round by a fixed stride (kernels.cuh RoundBase::dynamic)
    uint32_t gkr_linear_min = 9;         // ZKSC_GKR_LINEAR_MIN: zksc_gkr_prove's layers with at least this many bits per input label use the two-phase form
    bool tail_prelaunch = true;          // ZKSC_NO_PRELAUNCH=1: the resident kernel is launched in its first round, not behind the launch of the round before
    bool tail_dynamic = true;            // ZKSC_RES_STATIC=1: the resident kernel splits every round by a fixed stride (no work counter)
    unsigned long long tail_work = kTailWorkDefault;   // start threshold of the resident kernel (ZKSC_TAIL_WORK overrides, experiments)
    int res_occ[kResMaxDegree + 1] = {};       // resident CTAs per SM of resident_kernel<dsel>
    volatile uint64_t* tail_mail = nullptr;    // [tail_proofs_cap][kMailUnits]   {word | seq << 32}
    volatile uint64_t* tail_res = nullptr;     // [tail_units_cap]                {limb | seq << 32}
    unsigned long long* tail_mail_dev = nullptr;
    unsigned long long* tail_res_dev = nullptr;
    unsigned long long* tail_relay = nullptr;  // HBM [2][tail_proofs_cap][kMailUnits] units, then [2][tail_proofs_cap] tags (32-bit)
    Fr* tail_partials = nullptr;               // HBM [tail_part_cap] elements
    unsigned int* tail_counters = nullptr;     // HBM [tail_groups_cap]
    size_t tail_proofs_cap = 0, tail_units_cap = 0, tail_groups_cap = 0, tail_part_cap = 0;
    size_t tail_status_off = 0;                // the per-proof status units sit behind the result units in tail_res
    unsigned int tail_seq = 0;           // last sequence number handed out
    struct zksc_tables* active_tail = nullptr;   // the handle whose tail kernel is resident on `stream` (at most one)
    // ZKSC_PROFILE=1: host-side wall-clock split of every round of zksc_prove, printed to stderr (ns)
    bool profile = false;
    double prof_launch = 0, prof_wait = 0;
    // peer-memory exchange of the per-round partials (sharded contexts; kernels.cuh XchArgs)
    bool p2p = false;
    unsigned char* xch_local = nullptr;            // tags[2][n_ranks] (first kXchTagBytes) then data[2][n_ranks][kXchCap]
    void* xch_peer[kMaxRanks] = {};                // the same buffer of every rank (own entry = xch_local)
    unsigned int xch_seq = 0;
#if ZKSC_HAVE_NCCL_H
    ncclComm_t comm = nullptr;
#endif
    // measurement (bench.py): kernels launched so far, and optional per-launch CUDA-event timing
    unsigned long long launches = 0;
    bool timing = false;
    struct LaunchRec { cudaEvent_t e0, e1; unsigned int degree, fold; unsigned long long pairs, proofs; };
    std::vector<LaunchRec> recs;       // used entries: [0, n_recs)
    size_t n_recs = 0;
    std::vector<double> round_us;      // host wall time of every round of the latest zksc_prove (evaluations + transcript + bind)
};

struct zksc_tables {
    zksc_ctx* ctx;
    uint32_t n_vars, B, P, Dtot, E;
    uint32_t deg[ZKSC_MAX_PRODUCTS], koff[ZKSC_MAX_PRODUCTS], eoff[ZKSC_MAX_PRODUCTS];
    uint64_t n_local0;       // entries per table held by this rank at round 0
    Fr* orig = nullptr;      // [B][Dtot][n_local0]
    Fr* work = nullptr;      // [B][Dtot][max(n_local0/2,1)]
    Fr* tail = nullptr;      // [B][Dtot][n_ranks]   gathered residual (sharded contexts)
    // state
    uint32_t vars_left;
    uint64_t cur_n;          // entries per table in the current buffer
    int where;               // 0 orig, 1 work, 2 tail
    bool pending;
    std::vector<Fr> pending_chal;
    std::vector<FoldTab> pending_tab;   // shift tables of pending_chal (fr.cuh mul_fixed_rows)
    // round-0 evaluations computed by zksc_poly_sum, handed to the next round-0 zksc_round_evals once
    // (calculate_poly_sum followed by prove is the reference's calling pattern; the input is immutable)
    std::vector<uint64_t> r0_cache;
    bool in_prove = false;              // inside a prover's round loop: the next call is this handle's bind (what a queued-ahead resident kernel relies on)
    bool r0_valid = false;
    // The evaluations returned by the latest zksc_round_evals (valid until the next bind) and, after a
    // bind that followed them, the per-product claims h_p(r) = h_p,next(0) + h_p,next(1): with a claim the
    // next round's kernel skips evaluation point 1 and the host fills in h(1) = claim - h(0).
    std::vector<uint64_t> last_evals;   // [B][E][4]
    bool last_evals_valid = false;
    std::vector<FrH> claim;             // [B][P]
    bool claim_valid = false;
    // persistent tail kernel state
    bool tail_running = false;          // the kernel is resident and has rounds left
    bool tail_posted = false;           // a challenge has been posted whose round result has not been collected
    unsigned int tail_left = 0;         // rounds whose result has not been collected
    unsigned int tail_cur = 0;          // sequence number of the round posted last
    unsigned int tail_done = 0;         // rounds of the running resident kernel whose result has been collected
    std::chrono::steady_clock::time_point tail_seen;   // when the latest round's results were collected (the kernel has been waiting since)
    unsigned int tail_gather_round = kNoGather;   // sharded: its round (counted from its first) after which the shards are gathered
    unsigned long long tail_gather_local = 0;     // ... entries per local table at that point
    uint64_t tail_stride = 0;           // elements per table of `tail`
    bool copy_pending = false;          // zksc_tables_reupload_begin without its _end: `orig` is being written by the copy stream
    // single-process multi-GPU: the parent handle owns one child handle per device and no device memory of its own; a child returns
    // raw per-shard sums (no claim arithmetic, no point conversion: the parent does that on the totals)
    std::vector<zksc_tables*> kids;
    bool last_partial = false;          // the latest round's evaluations cover this rank's shard only (they still have to be added up)
    bool raw = false;                   // child of a multi-GPU handle
    bool raw_claim_next = false;        // ... whether the parent holds a claim for the round after the bind being applied
};

struct GpuPoolSession {     // the per-device worker threads spin while a multi-GPU call is in progress, and only then
    zksc_ctx* c;
    bool mine;
    explicit GpuPoolSession(zksc_ctx* ctx) : c(ctx), mine(false) {
        if (c && c->gpu_pool && !c->gpu_pool_active) { c->gpu_pool->begin(); c->gpu_pool_active = true; mine = true; }
    }
    ~GpuPoolSession() { if (mine) { c->gpu_pool_active = false; c->gpu_pool->end(); } }
};

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
            return (e_ == cudaErrorMemoryAllocation) ? ZKSC_ERR_OOM : ZKSC_ERR_CUDA;                      \
        }                                                                                                 \
    } while (0)
#define FAIL(code, msg)       \
    do {                      \
        ctx->err = (msg);     \
        return (code);        \
    } while (0)
#define TRY(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_ != ZKSC_OK) return rc_; \
    } while (0)

static int tail_stop(zksc_tables* t);
static int quiesce(zksc_ctx* ctx);
static int multi_tables_synth(zksc_ctx* ctx, uint32_t n_vars, uint32_t B, uint32_t P, const uint32_t* degree, uint64_t seed, zksc_tables** out);
static int multi_tables_upload(zksc_ctx* ctx, uint32_t n_vars, uint32_t B, uint32_t P, const uint32_t* degree, const uint64_t* const* host_tables, zksc_tables** out);
static int multi_tables_reset(zksc_tables* t);
static int multi_tables_free(zksc_tables* t);
static int multi_residual(zksc_tables* t, uint64_t* out);
#define NOT_MULTI(ctx_, what)                                                                                                  \
    do {                                                                                                                       \
        if ((ctx_) && !(ctx_)->kids.empty()) { (ctx_)->err = what ": not available on a multi-GPU context"; return ZKSC_ERR_UNSUPPORTED; } \
    } while (0)

static inline FrH to_host(const Fr& f) { FrH h; memcpy(h.v, f.l, 32); return h; }
static inline FrH load_h(const uint64_t* p) { FrH h; memcpy(h.v, p, 32); return h; }
static inline void store_h(uint64_t* p, const FrH& h) { memcpy(p, h.v, 32); }

// ------------------------------------------------------------------------------------------------
// library / context
// ------------------------------------------------------------------------------------------------
extern "C" const char* zksc_version(void) { return "zksc 0.1 (sm_100a)"; }

extern "C" int zksc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// per-degree launchers live in round_inst.cu (one object per degree, compiled in parallel)
#define ZKSC_DECL_ROUND(D)                                                                               \
    void zksc_launch_round_##D(int variant, bool staged, dim3 grid, cudaStream_t s, const RoundArgsT<1>& a);          \
    void zksc_launch_round_##D(int variant, bool staged, dim3 grid, cudaStream_t s, const RoundArgsT<kMaxBatch>& a);  \
    void zksc_prepare_round_##D();                                                                        \
    int zksc_occ_round_##D(int variant);
ZKSC_DECL_ROUND(1) ZKSC_DECL_ROUND(2) ZKSC_DECL_ROUND(3) ZKSC_DECL_ROUND(4)
ZKSC_DECL_ROUND(5) ZKSC_DECL_ROUND(6) ZKSC_DECL_ROUND(7) ZKSC_DECL_ROUND(8)

extern "C" int zksc_ctx_create(int device, zksc_ctx** out) {
    if (!out) return ZKSC_ERR_SHAPE;
    *out = nullptr;
    int n = zksc_device_count();
    if (n <= 0 || device < 0 || device >= n) {
        g_create_error = "no usable CUDA device (zksc has no CPU fallback)";
        return ZKSC_ERR_NO_DEVICE;
    }
    zksc_ctx* ctx = new zksc_ctx();
    ctx->device = device;
    auto fail = [&](cudaError_t e, const char* what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
        delete ctx;
        return ZKSC_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    ctx->sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    {   // keep up to 2 GiB of released blocks in the device's stream-ordered pool (dev_alloc below); larger ones go back at the next sync
        cudaMemPool_t pool;
        unsigned long long keep = 2ull << 30;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        cudaGetLastError();
    }
    if ((e = cudaMalloc(&ctx->counters, kRoundCounterWords * sizeof(unsigned int))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMemset(ctx->counters, 0, kRoundCounterWords * sizeof(unsigned int))) != cudaSuccess) return fail(e, "cudaMemset");
#define ZKSC_OCC(D) zksc_prepare_round_##D(); for (int v = 0; v < 6; v++) ctx->occ[D][v] = zksc_occ_round_##D(v);
    ZKSC_OCC(1) ZKSC_OCC(2) ZKSC_OCC(3) ZKSC_OCC(4) ZKSC_OCC(5) ZKSC_OCC(6) ZKSC_OCC(7) ZKSC_OCC(8)
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "occupancy query (is this an sm_100a device?)");
    {
        void* fh = nullptr;
        if ((e = cudaHostAlloc(&fh, 64, cudaHostAllocMapped)) != cudaSuccess) return fail(e, "cudaHostAlloc(flag)");
        memset(fh, 0, 64);
        ctx->flag_host = (volatile unsigned int*)fh;
        void* fd = nullptr;
        if ((e = cudaHostGetDevicePointer(&fd, fh, 0)) != cudaSuccess) return fail(e, "cudaHostGetDevicePointer(flag)");
        ctx->flag_dev = (unsigned int*)fd;
    }
    { const char* e_ = getenv("ZKSC_NO_MAPPED"); ctx->mapped_results = !(e_ && e_[0] == '1'); }
    { const char* e_ = getenv("ZKSC_PROFILE"); ctx->profile = (e_ && e_[0] == '1'); }
    { const char* e_ = getenv("ZKSC_NO_TAIL"); ctx->tail_enabled = !(e_ && e_[0] == '1'); }
    { const char* e_ = getenv("ZKSC_RES_STATIC"); ctx->tail_dynamic = !(e_ && e_[0] == '1'); }
    { const char* e_ = getenv("ZKSC_NO_PRELAUNCH"); ctx->tail_prelaunch = !(e_ && e_[0] == '1'); }
    { const char* e_ = getenv("ZKSC_GKR_LINEAR_MIN"); if (e_) ctx->gkr_linear_min = (uint32_t)strtoul(e_, nullptr, 10); }
    { const char* e_ = getenv("ZKSC_ROUND_STATIC"); ctx->round_dynamic = !(e_ && e_[0] == '1'); }
    { const char* e_ = getenv("ZKSC_DYN_MAX_GROUPS"); if (e_) ctx->dyn_max_groups = std::min<unsigned int>(kDynMaxGroups, (unsigned int)strtoul(e_, nullptr, 10)); }
    { const char* e_ = getenv("ZKSC_NO_FUSE"); ctx->fuse_products = !(e_ && e_[0] == '1'); }
    { const char* e_ = getenv("ZKSC_TAIL_WORK"); if (e_ && atoll(e_) > 0) ctx->tail_work = (unsigned long long)atoll(e_); }
    { const char* e_ = getenv("ZKSC_NO_STAGED"); ctx->staged = !(e_ && e_[0] == '1'); }
    { const char* e_ = getenv("ZKSC_STAGED_FOLD"); ctx->staged_fold = (e_ && e_[0] == '1'); }
    for (int d = 0; d <= kResMaxDegree; d++) ctx->res_occ[d] = zksc_resident_occ(d);
    {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (!coop) ctx->tail_enabled = false;     // the resident kernel's CTAs wait for each other: only with guaranteed co-residency
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "occupancy query (resident kernel)");
    *out = ctx;
    return ZKSC_OK;
}

extern "C" int zksc_ctx_destroy(zksc_ctx* ctx) {
    if (!ctx) return ZKSC_OK;
    if (ctx->kids.empty()) {
        cudaSetDevice(ctx->device);
        quiesce(ctx);
        cudaStreamSynchronize(ctx->stream);
    }
    if (!ctx->kids.empty()) {
        delete ctx->gpu_pool;
        for (zksc_ctx* k : ctx->kids) zksc_ctx_destroy(k);
        delete ctx->pool;
        delete ctx;
        return ZKSC_OK;
    }
    if (!ctx->host_reduce)       // (a multi-GPU child's peers live in the same process: plain pointers, nothing to close)
        for (int g = 0; g < ctx->n_ranks && g < kMaxRanks; g++)
            if (ctx->xch_peer[g] && g != ctx->rank) cudaIpcCloseMemHandle(ctx->xch_peer[g]);
    cudaFree(ctx->xch_local);
#if ZKSC_HAVE_NCCL_H
    if (ctx->comm) g_nccl.CommDestroy(ctx->comm);
#endif
    for (auto& r : ctx->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    cudaFree(ctx->partials);
    cudaFree(ctx->counters);
    cudaFree(ctx->results_dev);
    cudaFree(ctx->results_send);
    cudaFreeHost(ctx->results_host);
    cudaFreeHost((void*)ctx->flag_host);
    cudaFreeHost((void*)ctx->tail_mail);
    cudaFreeHost((void*)ctx->tail_res);
    cudaFree(ctx->tail_relay);
    cudaFree(ctx->tail_partials);
    cudaFree(ctx->tail_counters);
    delete ctx->pool;
    if (ctx->gkr_stage) cudaFreeHost(ctx->gkr_stage);
    if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->copy_event); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ZKSC_OK;
}

extern "C" const char* zksc_last_error(const zksc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int zksc_ctx_synchronize(zksc_ctx* ctx) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) {
        for (zksc_ctx* k : ctx->kids) { const int rc = zksc_ctx_synchronize(k); if (rc != ZKSC_OK) { ctx->err = k->err; return rc; } }
        return ZKSC_OK;
    }
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

extern "C" int zksc_ctx_rank(const zksc_ctx* ctx, int* rank, int* n_ranks) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (rank) *rank = ctx->rank;
    if (n_ranks) *n_ranks = ctx->n_ranks;
    return ZKSC_OK;
}

extern "C" int zksc_ctx_peer_exchange(const zksc_ctx* ctx) { return (ctx && ctx->p2p) ? 1 : 0; }

extern "C" uint64_t zksc_ctx_gather_entries(const zksc_ctx* ctx) { return ctx ? (uint64_t)ctx->gather_entries : 0; }

extern "C" void* zksc_ctx_stream(zksc_ctx* ctx) { return ctx ? (void*)(ctx->kids.empty() ? ctx->stream : ctx->kids[0]->stream) : nullptr; }

extern "C" unsigned long long zksc_ctx_launch_count(const zksc_ctx* ctx) {
    if (!ctx) return 0;
    unsigned long long n = ctx->launches;
    for (const zksc_ctx* k : ctx->kids) n += k->launches;
    return n;
}

extern "C" int zksc_ctx_timing(zksc_ctx* ctx, int enable) {
    if (!ctx) return ZKSC_ERR_STATE;
    NOT_MULTI(ctx, "zksc_ctx_timing");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->timing = enable != 0;
    ctx->n_recs = 0;
    return ZKSC_OK;
}

extern "C" int zksc_ctx_timing_read(zksc_ctx* ctx, uint32_t cap, uint32_t* n_out, float* ms, uint32_t* degree, uint32_t* fold, uint64_t* pairs,
                                    uint64_t* proofs) {
    if (!ctx || !n_out) return ZKSC_ERR_STATE;
    NOT_MULTI(ctx, "zksc_ctx_timing_read");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    uint32_t n = 0;
    for (size_t i = 0; i < ctx->n_recs && n < cap; i++, n++) {
        const auto& r = ctx->recs[i];
        if (ms) CK(cudaEventElapsedTime(ms + n, r.e0, r.e1));
        if (degree) degree[n] = r.degree;
        if (fold) fold[n] = r.fold;
        if (pairs) pairs[n] = r.pairs;
        if (proofs) proofs[n] = r.proofs;
    }
    *n_out = n;
    ctx->n_recs = 0;
    return ZKSC_OK;
}

extern "C" int zksc_ctx_round_times(const zksc_ctx* ctx, uint32_t cap, uint32_t* n_out, double* us) {
    if (!ctx || !n_out) return ZKSC_ERR_STATE;
    uint32_t n = 0;
    for (; n < cap && n < ctx->round_us.size(); n++)
        if (us) us[n] = ctx->round_us[n];
    *n_out = n;
    return ZKSC_OK;
}

// The integer roof of this device, measured: 8 independent 32 x 32 -> 64 multiply-adds (IMAD.WIDE.U32, the instruction every limb
// product of fr.cuh compiles to) per thread and iteration, nothing else in the loop, all SMs full.
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed;
    uint32_t c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = a + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(c[2 * i]), "+r"(c[2 * i + 1]) : "r"(c[(2 * i + 3) & 15]), "r"(c[(2 * i + 6) & 15]));
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) x ^= c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
extern "C" int zksc_int_peak(zksc_ctx* ctx, double* limb_products_per_second) {
    if (!ctx || !limb_products_per_second) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) ctx = ctx->kids[0];
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    const int blocks = ctx->sms * 8, iters = 4096;
    uint32_t* buf = nullptr;
    CK(cudaMallocAsync((void**)&buf, (size_t)blocks * 256 * sizeof(uint32_t), ctx->stream));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {          // the first one warms up
        CK(cudaEventRecord(e0, ctx->stream));
        int_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, iters, 1u);
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    ctx->launches += 6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFreeAsync(buf, ctx->stream);
    CK(cudaGetLastError());
    *limb_products_per_second = (double)blocks * 256 * iters * 8 / (best * 1e-3);
    return ZKSC_OK;
}

// bracket one round-kernel launch with events when timing is on
static int timing_open(zksc_ctx* ctx, unsigned int degree, bool fold, unsigned long long pairs, unsigned long long proofs) {
    if (!ctx->timing) return ZKSC_OK;
    if (ctx->n_recs == ctx->recs.size()) {
        zksc_ctx::LaunchRec r{};
        CK(cudaEventCreate(&r.e0));
        CK(cudaEventCreate(&r.e1));
        ctx->recs.push_back(r);
    }
    auto& r = ctx->recs[ctx->n_recs];
    r.degree = degree; r.fold = fold ? 1u : 0u; r.pairs = pairs; r.proofs = proofs;
    CK(cudaEventRecord(r.e0, ctx->stream));
    return ZKSC_OK;
}
static int timing_close(zksc_ctx* ctx) {
    if (!ctx->timing) return ZKSC_OK;
    CK(cudaEventRecord(ctx->recs[ctx->n_recs].e1, ctx->stream));
    ctx->n_recs++;
    return ZKSC_OK;
}

extern "C" int zksc_comm_unique_id(uint8_t out_id[128]) {
#if ZKSC_HAVE_NCCL_H
    std::string err;
    if (!g_nccl.load(err)) { g_create_error = err; return ZKSC_ERR_COMM; }
    ncclUniqueId id;
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return ZKSC_ERR_COMM; }
    memcpy(out_id, &id, 128);
    return ZKSC_OK;
#else
    g_create_error = "built without nccl.h";
    return ZKSC_ERR_COMM;
#endif
}

#if ZKSC_HAVE_NCCL_H
// Open every rank's exchange buffer through CUDA IPC (handles travel once through ncclAllGather).  All ranks
// must agree on whether the peer path is usable: a last all-gather of one status word settles it; on any
// failure every rank keeps the ncclAllGather + host-sum path.  ZKSC_NO_P2P=1 (on every rank) forces that path.
static int setup_peer_exchange(zksc_ctx* ctx) {
    ctx->p2p = false;
    const int G = ctx->n_ranks;
    { const char* e_ = getenv("ZKSC_NO_P2P"); if (e_ && e_[0] == '1') return ZKSC_OK; }
    if (G > kMaxRanks) return ZKSC_OK;
    const size_t bytes = kXchStageOffset(G) + kXchStageElems * sizeof(Fr);
    struct Msg { cudaIpcMemHandle_t h; unsigned int ok; unsigned int pad[15]; };
    static_assert(sizeof(Msg) == 128, "exchange handle message");
    Msg mine;
    memset(&mine, 0, sizeof(mine));
    mine.ok = 1;
    if (cudaMalloc(&ctx->xch_local, bytes) != cudaSuccess || cudaMemsetAsync(ctx->xch_local, 0, bytes, ctx->stream) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.h, ctx->xch_local) != cudaSuccess) {
        cudaGetLastError();
        mine.ok = 0;
    }
    unsigned char* dev = nullptr;
    CK(cudaMalloc(&dev, sizeof(Msg) * (G + 1)));
    std::vector<Msg> all(G);
    auto gather = [&](const Msg& m) -> int {
        CK(cudaMemcpyAsync(dev + sizeof(Msg) * G, &m, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream));
        ncclResult_t r = g_nccl.AllGather(dev + sizeof(Msg) * G, dev, sizeof(Msg), ncclUint8, ctx->comm, ctx->stream);
        if (r != ncclSuccess) FAIL(ZKSC_ERR_COMM, std::string("ncclAllGather(ipc handles): ") + g_nccl.GetErrorString(r));
        CK(cudaMemcpyAsync(all.data(), dev, sizeof(Msg) * G, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return ZKSC_OK;
    };
    int rc = gather(mine);
    if (rc != ZKSC_OK) { cudaFree(dev); return rc; }
    bool ok = true;
    for (int g = 0; g < G; g++) ok = ok && all[g].ok;
    if (ok) {
        for (int g = 0; g < G && ok; g++) {
            if (g == ctx->rank) { ctx->xch_peer[g] = ctx->xch_local; continue; }
            if (cudaIpcOpenMemHandle(&ctx->xch_peer[g], all[g].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ctx->xch_peer[g] = nullptr;
                ok = false;
            }
        }
    }
    mine.ok = ok ? 1 : 0;
    rc = gather(mine);                                  // second pass: did everyone manage to open everyone?
    cudaFree(dev);
    if (rc != ZKSC_OK) return rc;
    for (int g = 0; g < G; g++) ok = ok && all[g].ok;
    ctx->p2p = ok;
    return ZKSC_OK;
}
#endif

extern "C" int zksc_comm_init(zksc_ctx* ctx, int n_ranks, int rank, const uint8_t unique_id[128]) {
    if (!ctx) return ZKSC_ERR_STATE;
    NOT_MULTI(ctx, "zksc_comm_init");
    if (ctx->host_reduce) FAIL(ZKSC_ERR_STATE, "a child of a multi-GPU context has no communicator");
    if (n_ranks < 1 || (n_ranks & (n_ranks - 1)) || rank < 0 || rank >= n_ranks) FAIL(ZKSC_ERR_SHAPE, "n_ranks must be a power of two and 0 <= rank < n_ranks");
    TRY(quiesce(ctx));
    if (n_ranks == 1) { ctx->rank = 0; ctx->n_ranks = 1; return ZKSC_OK; }
#if ZKSC_HAVE_NCCL_H
    if (!g_nccl.load(ctx->err)) return ZKSC_ERR_COMM;
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, n_ranks, id, rank);
    if (r != ncclSuccess) FAIL(ZKSC_ERR_COMM, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    ctx->rank = rank;
    ctx->n_ranks = n_ranks;
    if (!getenv("ZKSC_TAIL_WORK")) ctx->tail_work = kTailWorkSharded;
    {
        unsigned long long g = kGatherDefault;
        const char* e_ = getenv("ZKSC_GATHER_ENTRIES");
        if (e_ && atoll(e_) > 0) g = (unsigned long long)atoll(e_);
        while (g & (g - 1)) g &= g - 1;                     // a power of two ...
        if (g < 2ull * n_ranks) g = 2ull * n_ranks;         // ... with at least one pair per rank left
        ctx->gather_entries = g;
    }
    // result buffers are sized per rank count: drop them so the next use re-allocates
    cudaFree(ctx->results_dev); cudaFree(ctx->results_send); cudaFreeHost(ctx->results_host);
    ctx->results_dev = nullptr; ctx->results_send = nullptr; ctx->results_host = nullptr; ctx->results_cap = 0;
    return setup_peer_exchange(ctx);
#else
    FAIL(ZKSC_ERR_COMM, "built without nccl.h");
#endif
}

static int ensure_results(zksc_ctx* ctx, size_t elems) {
    if (elems <= ctx->results_cap) return ZKSC_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->results_dev); cudaFree(ctx->results_send); cudaFreeHost(ctx->results_host);
    ctx->results_dev = nullptr; ctx->results_send = nullptr; ctx->results_host = nullptr; ctx->results_cap = 0;
    CK(cudaMalloc(&ctx->results_dev, elems * ctx->n_ranks * sizeof(Fr)));
    CK(cudaMalloc(&ctx->results_send, elems * sizeof(Fr)));
    CK(cudaHostAlloc(&ctx->results_host, elems * ctx->n_ranks * sizeof(Fr), cudaHostAllocMapped));
    { void* d = nullptr; CK(cudaHostGetDevicePointer(&d, ctx->results_host, 0)); ctx->results_host_dev = (Fr*)d; }
    ctx->results_cap = elems;
    return ZKSC_OK;
}
static int ensure_partials(zksc_ctx* ctx, size_t elems) {
    if (elems <= ctx->partials_cap) return ZKSC_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->partials);
    ctx->partials = nullptr; ctx->partials_cap = 0;
    CK(cudaMalloc(&ctx->partials, elems * sizeof(Fr)));
    ctx->partials_cap = elems;
    return ZKSC_OK;
}

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync on the context's stream): an allocation or a
// release costs microseconds and never synchronises the device -- the GKR driver builds and drops four tables per layer,
// and a plain cudaFree would wait for a resident rounds kernel to time out.
static cudaError_t dev_alloc(zksc_ctx* ctx, void** p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream); }
static void dev_free(zksc_ctx* ctx, void* p) { if (p) cudaFreeAsync(p, ctx->stream); }
struct DevBuf {
    zksc_ctx* ctx;
    Fr* p = nullptr;
    explicit DevBuf(zksc_ctx* c) : ctx(c) {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { dev_free(ctx, p); }
};

// ------------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------------
static int tables_alloc(zksc_ctx* ctx, uint32_t n_vars, uint32_t B, uint32_t P, const uint32_t* degree, zksc_tables** out) {
    if (!ctx || !out || !degree) return ZKSC_ERR_SHAPE;
    *out = nullptr;
    TRY(quiesce(ctx));
    if (B < 1 || P < 1 || P > ZKSC_MAX_PRODUCTS) FAIL(ZKSC_ERR_SHAPE, "need 1 <= n_products <= ZKSC_MAX_PRODUCTS and n_proofs >= 1");
    if (n_vars > 40) FAIL(ZKSC_ERR_SHAPE, "n_vars too large");
    int lg = 0;
    while ((1 << lg) < ctx->n_ranks) lg++;
    if ((int)n_vars < lg) FAIL(ZKSC_ERR_SHAPE, "tables have fewer entries than there are ranks");
    zksc_tables* t = new zksc_tables();
    t->ctx = ctx; t->n_vars = n_vars; t->B = B; t->P = P; t->Dtot = 0; t->E = 0;
    for (uint32_t p = 0; p < P; p++) {
        if (degree[p] < 1 || degree[p] > ZKSC_MAX_DEGREE) { delete t; FAIL(ZKSC_ERR_UNSUPPORTED, "product degree must be in 1..ZKSC_MAX_DEGREE"); }
        t->deg[p] = degree[p]; t->koff[p] = t->Dtot; t->eoff[p] = t->E;
        t->Dtot += degree[p]; t->E += degree[p] + 1;
    }
    t->n_local0 = (1ull << n_vars) / ctx->n_ranks;
    CK(cudaSetDevice(ctx->device));
    size_t n_orig = (size_t)B * t->Dtot * t->n_local0;
    size_t n_work = (size_t)B * t->Dtot * (t->n_local0 > 1 ? t->n_local0 / 2 : 1);
    cudaError_t e = dev_alloc(ctx, (void**)&t->orig, n_orig * sizeof(Fr));
    if (e == cudaSuccess) e = dev_alloc(ctx, (void**)&t->work, n_work * sizeof(Fr));
    if (ctx->n_ranks > 1) t->tail_stride = std::max<uint64_t>(ctx->gather_entries, (uint64_t)ctx->n_ranks);
    if (e == cudaSuccess && ctx->n_ranks > 1) e = dev_alloc(ctx, (void**)&t->tail, (size_t)B * t->Dtot * t->tail_stride * sizeof(Fr));
    if (e != cudaSuccess) {
        dev_free(ctx, t->orig); dev_free(ctx, t->work); dev_free(ctx, t->tail);
        delete t;
        ctx->err = std::string("cudaMalloc(tables): ") + cudaGetErrorString(e);
        cudaGetLastError();
        return ZKSC_ERR_OOM;
    }
    {
        size_t need = (size_t)B * t->E, need2 = (size_t)B * t->Dtot;
        int rc = ensure_results(ctx, need > need2 ? need : need2);
        if (rc != ZKSC_OK) { dev_free(ctx, t->orig); dev_free(ctx, t->work); dev_free(ctx, t->tail); delete t; return rc; }
    }
    zksc_tables_reset(t);
    *out = t;
    return ZKSC_OK;
}

static int copy_finish(zksc_tables* t);
extern "C" int zksc_tables_reset(zksc_tables* t) {
    if (!t) return ZKSC_ERR_STATE;
    if (!t->kids.empty()) return multi_tables_reset(t);
    TRY(copy_finish(t));
    TRY(tail_stop(t));
    t->vars_left = t->n_vars;
    t->cur_n = t->n_local0;
    t->where = 0;
    t->pending = false;
    t->pending_chal.clear();
    t->last_evals_valid = false;
    t->claim_valid = false;
    return ZKSC_OK;
}

extern "C" int zksc_tables_free(zksc_tables* t) {
    if (!t) return ZKSC_OK;
    if (!t->kids.empty()) return multi_tables_free(t);
    zksc_ctx* ctx = t->ctx;
    cudaSetDevice(ctx->device);
    copy_finish(t);
    tail_stop(t);
    dev_free(ctx, t->orig); dev_free(ctx, t->work); dev_free(ctx, t->tail);   // stream-ordered: after everything queued so far
    delete t;
    return ZKSC_OK;
}

extern "C" int zksc_tables_vars_left(const zksc_tables* t, uint32_t* out) {   // (a multi-GPU parent keeps its own count)
    if (!t || !out) return ZKSC_ERR_STATE;
    *out = t->vars_left;
    return ZKSC_OK;
}

__global__ void __launch_bounds__(256) pick_shard_kernel(const Fr* full, Fr* out, unsigned long long n_local, unsigned long long G, unsigned long long g) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += stride) st256(out + i, ld256(full + i * G + g));
}

static inline int grid_for(const zksc_ctx* ctx, unsigned long long n, int threads, int per_sm) {
    unsigned long long blocks = (n + threads - 1) / threads;
    unsigned long long cap = (unsigned long long)ctx->sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// copy (and, on a sharded context, stride-pick) the caller's full tables into t->orig
static int copy_finish(zksc_tables* t);
static int upload_into(zksc_tables* t, const uint64_t* const* host_tables, bool local) {
    zksc_ctx* ctx = t->ctx;
    CK(cudaSetDevice(ctx->device));
    TRY(copy_finish(t));
    TRY(quiesce(ctx));
    t->r0_valid = false;
    const uint64_t N = 1ull << t->n_vars;
    const size_t n_tabs = (size_t)t->B * t->Dtot;
    if (ctx->n_ranks == 1 || local) {
        const uint64_t NL = t->n_local0;
        for (size_t i = 0; i < n_tabs; i++) CK(cudaMemcpyAsync(t->orig + i * NL, host_tables[i], NL * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // stage one full table at a time, keep the entries with index = rank (mod n_ranks)
        DevBuf stage(ctx);
        CK(dev_alloc(ctx, (void**)&stage.p, N * sizeof(Fr)));
        for (size_t i = 0; i < n_tabs; i++) {
            CK(cudaMemcpyAsync(stage.p, host_tables[i], N * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
            pick_shard_kernel<<<grid_for(ctx, t->n_local0, 256, 8), 256, 0, ctx->stream>>>(stage.p, t->orig + i * t->n_local0, t->n_local0, ctx->n_ranks, ctx->rank);
            ctx->launches++;
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));  // host buffers are only borrowed for the call
    return ZKSC_OK;
}

extern "C" int zksc_tables_upload(zksc_ctx* ctx, uint32_t n_vars, uint32_t n_proofs, uint32_t n_products, const uint32_t* degree,
                                  const uint64_t* const* host_tables, zksc_tables** out) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!host_tables) FAIL(ZKSC_ERR_SHAPE, "host_tables is NULL");
    if (!ctx->kids.empty()) {
        if (!out) return ZKSC_ERR_SHAPE;
        int lg = 0;
        while ((1u << lg) < ctx->kids.size()) lg++;
        if ((int)n_vars < lg || n_vars > 40) FAIL(ZKSC_ERR_SHAPE, "tables have fewer entries than there are devices (or n_vars is too large)");
        return multi_tables_upload(ctx, n_vars, n_proofs, n_products, degree, host_tables, out);
    }
    zksc_tables* t = nullptr;
    TRY(tables_alloc(ctx, n_vars, n_proofs, n_products, degree, &t));
    int rc = upload_into(t, host_tables, false);
    if (rc != ZKSC_OK) { std::string keep = ctx->err; zksc_tables_free(t); ctx->err = keep; return rc; }
    *out = t;
    return ZKSC_OK;
}

extern "C" int zksc_tables_upload_local(zksc_ctx* ctx, uint32_t n_vars, uint32_t n_proofs, uint32_t n_products, const uint32_t* degree,
                                        const uint64_t* const* host_local_tables, zksc_tables** out) {
    if (!ctx) return ZKSC_ERR_STATE;
    NOT_MULTI(ctx, "zksc_tables_upload_local");
    if (!host_local_tables) FAIL(ZKSC_ERR_SHAPE, "host_local_tables is NULL");
    zksc_tables* t = nullptr;
    TRY(tables_alloc(ctx, n_vars, n_proofs, n_products, degree, &t));
    int rc = upload_into(t, host_local_tables, true);
    if (rc != ZKSC_OK) { std::string keep = ctx->err; zksc_tables_free(t); ctx->err = keep; return rc; }
    *out = t;
    return ZKSC_OK;
}

extern "C" int zksc_tables_reupload(zksc_tables* t, const uint64_t* const* host_tables, int local) {
    if (!t) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    NOT_MULTI(ctx, "zksc_tables_reupload");
    if (!host_tables) FAIL(ZKSC_ERR_SHAPE, "host_tables is NULL");
    zksc_tables_reset(t);
    return upload_into(t, host_tables, local != 0);
}

static int copy_finish(zksc_tables* t) { return t->copy_pending ? zksc_tables_reupload_end(t) : ZKSC_OK; }
// Double buffering across proofs: refill handle B on a second stream while handle A is being proved on the compute stream.
extern "C" int zksc_tables_reupload_begin(zksc_tables* t, const uint64_t* const* host_tables, int local) {
    if (!t) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (!host_tables) FAIL(ZKSC_ERR_SHAPE, "host_tables is NULL");
    NOT_MULTI(ctx, "zksc_tables_reupload_begin");
    if (t->copy_pending) FAIL(ZKSC_ERR_STATE, "a refill of this handle is already in flight; call zksc_tables_reupload_end first");
    TRY(zksc_tables_reset(t));      // stops this handle's resident kernel, if any
    TRY(quiesce(ctx));              // single-threaded API: another handle's resident kernel can only be in its exit phase here
    if (ctx->n_ranks > 1 && !local) return upload_into(t, host_tables, false);   // the strided shard pick needs the compute stream: synchronous
    CK(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
    }
    t->r0_valid = false;
    // everything queued on the compute stream so far (earlier uses of this handle, stream-ordered frees) comes first
    CK(cudaEventRecord(ctx->copy_event, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));
    const uint64_t NL = t->n_local0;
    const size_t n_tabs = (size_t)t->B * t->Dtot;
    for (size_t i = 0; i < n_tabs; i++) CK(cudaMemcpyAsync(t->orig + i * NL, host_tables[i], NL * sizeof(Fr), cudaMemcpyHostToDevice, ctx->copy_stream));
    t->copy_pending = true;
    return ZKSC_OK;
}

extern "C" int zksc_tables_reupload_end(zksc_tables* t) {
    if (!t) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (!t->copy_pending) return ZKSC_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->copy_stream));   // the host buffers are borrowed until here
    t->copy_pending = false;
    return ZKSC_OK;
}

extern "C" int zksc_tables_read_local(zksc_tables* t, uint64_t* out) {
    if (!t || !out) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    NOT_MULTI(ctx, "zksc_tables_read_local");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    const size_t n = (size_t)t->B * t->Dtot * t->n_local0;
    CK(cudaMemcpyAsync(out, t->orig, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

extern "C" int zksc_tables_synth(zksc_ctx* ctx, uint32_t n_vars, uint32_t n_proofs, uint32_t n_products, const uint32_t* degree, uint64_t seed,
                                 zksc_tables** out) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) {
        if (!out) return ZKSC_ERR_SHAPE;
        int lg = 0;
        while ((1u << lg) < ctx->kids.size()) lg++;
        if ((int)n_vars < lg || n_vars > 40) FAIL(ZKSC_ERR_SHAPE, "tables have fewer entries than there are devices (or n_vars is too large)");
        GpuPoolSession session(ctx);
        return multi_tables_synth(ctx, n_vars, n_proofs, n_products, degree, seed, out);
    }
    zksc_tables* t = nullptr;
    TRY(tables_alloc(ctx, n_vars, n_proofs, n_products, degree, &t));
    for (uint32_t b = 0; b < t->B; b++)
        for (uint32_t k = 0; k < t->Dtot; k++) {
            synth_kernel<<<grid_for(ctx, t->n_local0, 256, 8), 256, 0, ctx->stream>>>(t->orig + ((size_t)b * t->Dtot + k) * t->n_local0, t->n_local0,
                                                                                        seed + b, k, ctx->rank, ctx->n_ranks);
            ctx->launches++;
        }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { zksc_tables_free(t); ctx->err = std::string("synth: ") + cudaGetErrorString(e); return ZKSC_ERR_CUDA; }
    *out = t;
    return ZKSC_OK;
}

extern "C" int zksc_tables_alloc(zksc_ctx* ctx, uint32_t n_vars, uint32_t n_proofs, uint32_t n_products, const uint32_t* degree, zksc_tables** out) {
    if (!ctx) return ZKSC_ERR_STATE;
    NOT_MULTI(ctx, "zksc_tables_alloc");
    return tables_alloc(ctx, n_vars, n_proofs, n_products, degree, out);
}

// common checks of the fill calls: the handle must be unbound (its `orig` tables are being replaced)
static int fill_prepare(zksc_tables* t, uint32_t table) {
    zksc_ctx* ctx = t->ctx;
    NOT_MULTI(ctx, "zksc_tables_fill_*");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    if (table >= t->B * t->Dtot) FAIL(ZKSC_ERR_SHAPE, "table index out of range");
    TRY(zksc_tables_reset(t));
    t->r0_valid = false;
    return ZKSC_OK;
}

extern "C" int zksc_tables_fill_outer(zksc_tables* t, uint32_t table, int mul, const uint64_t* a, uint64_t na, const uint64_t* b, uint64_t nb) {
    if (!t) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (!a || !b || !na || !nb) FAIL(ZKSC_ERR_SHAPE, "empty operand");
    {
        unsigned long long prod = 0;
        if ((na & (na - 1)) || (nb & (nb - 1)) || __builtin_umulll_overflow(na, nb, &prod) || prod != (1ull << t->n_vars))
            FAIL(ZKSC_ERR_SHAPE, "add_distinct / mul_distinct: the operand sizes must be powers of two that multiply to 2^n_vars");
    }
    TRY(fill_prepare(t, table));
    DevBuf da(ctx), db(ctx);
    CK(dev_alloc(ctx, (void**)&da.p, na * sizeof(Fr)));
    CK(dev_alloc(ctx, (void**)&db.p, nb * sizeof(Fr)));
    CK(cudaMemcpyAsync(da.p, a, na * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(db.p, b, nb * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    outer_fill_kernel<<<grid_for(ctx, t->n_local0, 256, 8), 256, 0, ctx->stream>>>(mul, da.p, db.p, nb, t->orig + (size_t)table * t->n_local0, t->n_local0, ctx->rank,
                                                                                   ctx->n_ranks);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

extern "C" int zksc_tables_fill_sparse(zksc_tables* t, uint32_t table, const uint64_t* idx, const uint64_t* vals, uint64_t count) {
    if (!t) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (count && (!idx || !vals)) FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    for (uint64_t i = 0; i < count; i++)
        if (idx[i] >> t->n_vars) FAIL(ZKSC_ERR_SHAPE, "sparse entry index out of range");
    TRY(fill_prepare(t, table));
    Fr* dst = t->orig + (size_t)table * t->n_local0;
    CK(cudaMemsetAsync(dst, 0, t->n_local0 * sizeof(Fr), ctx->stream));
    if (count) {
        DevBuf dv(ctx), di(ctx);
        CK(dev_alloc(ctx, (void**)&dv.p, count * sizeof(Fr)));
        CK(dev_alloc(ctx, (void**)&di.p, (count + 3) / 4 * sizeof(Fr)));
        CK(cudaMemcpyAsync(dv.p, vals, count * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(di.p, idx, count * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        scatter_kernel<<<(unsigned int)((count + 255) / 256), 256, 0, ctx->stream>>>((const unsigned long long*)di.p, dv.p, count, dst, ctx->rank, ctx->n_ranks);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return ZKSC_OK;
}

extern "C" int zksc_tables_fill_dense(zksc_tables* t, uint32_t table, const uint64_t* evals) {
    if (!t) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (!evals) FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    TRY(fill_prepare(t, table));
    Fr* dst = t->orig + (size_t)table * t->n_local0;
    if (ctx->n_ranks == 1) {
        CK(cudaMemcpyAsync(dst, evals, t->n_local0 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        const uint64_t N = 1ull << t->n_vars;
        DevBuf stage(ctx);
        CK(dev_alloc(ctx, (void**)&stage.p, N * sizeof(Fr)));
        CK(cudaMemcpyAsync(stage.p, evals, N * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
        pick_shard_kernel<<<grid_for(ctx, t->n_local0, 256, 8), 256, 0, ctx->stream>>>(stage.p, dst, t->n_local0, ctx->n_ranks, ctx->rank);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

// current buffer geometry
struct Geo {
    Fr* base;
    unsigned long long tab_stride, proof_stride;
};
static Geo geo_of(const zksc_tables* t, int where) {
    Geo g;
    if (where == 0) { g.base = t->orig; g.tab_stride = t->n_local0; }
    else if (where == 1) { g.base = t->work; g.tab_stride = t->n_local0 > 1 ? t->n_local0 / 2 : 1; }
    else { g.base = t->tail; g.tab_stride = t->tail_stride; }
    g.proof_stride = g.tab_stride * t->Dtot;
    return g;
}


// ------------------------------------------------------------------------------------------------
// resident rounds kernel: host side (resident_kernel.cuh has the protocol)
// ------------------------------------------------------------------------------------------------
// tail_counters: [groups] arrival counters | (cache-line aligned) [groups][kResWarps][kWorkCtrWords] work counters
static size_t kTailWorkBase(size_t groups) { return (groups + kWorkCtrWords - 1) / kWorkCtrWords * kWorkCtrWords; }
static size_t kTailCounterWords(size_t groups) { return kTailWorkBase(groups) + groups * kResWarps * kWorkCtrWords; }
static int tail_ensure(zksc_ctx* ctx, size_t proofs, size_t units, size_t groups, size_t partials) {
    if (proofs <= ctx->tail_proofs_cap && units <= ctx->tail_units_cap && groups <= ctx->tail_groups_cap && partials <= ctx->tail_part_cap) return ZKSC_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFreeHost((void*)ctx->tail_mail); cudaFreeHost((void*)ctx->tail_res); cudaFree(ctx->tail_relay); cudaFree(ctx->tail_partials); cudaFree(ctx->tail_counters);
    ctx->tail_mail = nullptr; ctx->tail_res = nullptr; ctx->tail_relay = nullptr; ctx->tail_partials = nullptr; ctx->tail_counters = nullptr;
    ctx->tail_proofs_cap = ctx->tail_units_cap = ctx->tail_groups_cap = ctx->tail_part_cap = 0;
    proofs = std::max(proofs, ctx->tail_proofs_cap); units = std::max(units, ctx->tail_units_cap);
    void *m = nullptr, *r = nullptr, *d = nullptr;
    CK(cudaHostAlloc(&m, proofs * kMailUnits * 8, cudaHostAllocMapped));
    CK(cudaHostAlloc(&r, (units + proofs) * 8, cudaHostAllocMapped));
    memset(m, 0, proofs * kMailUnits * 8);
    memset(r, 0, (units + proofs) * 8);
    ctx->tail_status_off = units;
    ctx->tail_mail = (volatile uint64_t*)m; ctx->tail_res = (volatile uint64_t*)r;
    CK(cudaHostGetDevicePointer(&d, m, 0)); ctx->tail_mail_dev = (unsigned long long*)d;
    CK(cudaHostGetDevicePointer(&d, r, 0)); ctx->tail_res_dev = (unsigned long long*)d;
    const size_t relay_words = 2 * proofs * (kMailUnits + 1);
    CK(cudaMalloc(&ctx->tail_relay, relay_words * sizeof(unsigned long long)));
    CK(cudaMalloc(&ctx->tail_partials, partials * sizeof(Fr)));
    CK(cudaMalloc(&ctx->tail_counters, kTailCounterWords(groups) * sizeof(unsigned int)));     // arrival counters, then work counters
    CK(cudaMemsetAsync(ctx->tail_relay, 0, relay_words * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->tail_counters, 0, kTailCounterWords(groups) * sizeof(unsigned int), ctx->stream));
    ctx->tail_proofs_cap = proofs; ctx->tail_units_cap = units; ctx->tail_groups_cap = groups; ctx->tail_part_cap = partials;
    return ZKSC_OK;
}

// post the pending challenges' fold tables as round `seq` of every proof (words in FoldTabS order: fr.cuh foldtabs_index)
static void tail_post(zksc_tables* t, unsigned int seq) {
    zksc_ctx* ctx = t->ctx;
    for (uint32_t b = 0; b < t->B; b++) {
        const FoldTab& w = t->pending_tab[b];
        volatile uint64_t* m = ctx->tail_mail + (size_t)b * kMailUnits;
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) m[foldtabs_index(i, j)] = (uint64_t)w.w[i][j] | ((uint64_t)seq << 32);
    }
    t->tail_cur = seq;
    t->tail_posted = true;
}

constexpr int kTailExpired = 1;   // internal status of tail_wait (never crosses the C ABI)
// Forget a resident kernel that is gone or unusable and leave the context in a state from which ordinary launches work: the
// handle's and the context's bookkeeping is cleared whatever happened, the kernel (if it is still there) is told to leave, the
// stream is drained and every message buffer is wiped, so that no stale sequence number or failure tag survives.
static int tail_forget(zksc_tables* t, bool disable) {
    zksc_ctx* ctx = t->ctx;
    if (ctx->tail_mail)
        for (uint32_t b = 0; b < t->B && b < ctx->tail_proofs_cap; b++) ctx->tail_mail[(size_t)b * kMailUnits] = (uint64_t)kTailAbort << 32;
    t->tail_running = false; t->tail_posted = false; t->tail_left = 0;
    if (ctx->active_tail == t) ctx->active_tail = nullptr;
    if (disable) ctx->tail_enabled = false;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (ctx->tail_mail) memset((void*)ctx->tail_mail, 0, ctx->tail_proofs_cap * kMailUnits * 8);
    if (ctx->tail_res) memset((void*)ctx->tail_res, 0, (ctx->tail_units_cap + ctx->tail_proofs_cap) * 8);
    ctx->tail_seq = 0;
    if (e == cudaSuccess && ctx->tail_relay) e = cudaMemsetAsync(ctx->tail_relay, 0, 2 * ctx->tail_proofs_cap * (kMailUnits + 1) * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess && ctx->tail_counters) e = cudaMemsetAsync(ctx->tail_counters, 0, kTailCounterWords(ctx->tail_groups_cap) * sizeof(unsigned int), ctx->stream);
    if (e == cudaSuccess && ctx->xch_local) e = cudaMemsetAsync(ctx->xch_local + kXchUnitsOffset(ctx->n_ranks), 0, kXchUnitsBytes(ctx->n_ranks), ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("resident kernel clean-up: ") + cudaGetErrorString(e); return ZKSC_ERR_CUDA; }
    return ZKSC_OK;
}
// The kernel left on its own (mailbox timeout) before folding anything of the posted round: the pending challenge is still
// unapplied.  This happens when kernel launches are synchronous (under Nsight Compute every launch blocks until the kernel has
// ended, so the host can never answer a resident kernel) or when the host thread was stopped for a second: the context goes back
// to one launch per round for good.  The recovered round is evaluated in full -- point 1 included, not derived from the claim.
static int tail_expired(zksc_tables* t) {
    t->claim_valid = false;
    return tail_forget(t, true);
}
// wait for round `seq` of every (proof, product); copy the evaluations (all points but 1) to out when given
static int tail_wait(zksc_tables* t, unsigned int seq, uint64_t* out) {
    zksc_ctx* ctx = t->ctx;
    unsigned long long spins = 0;
    uint32_t timed_out = 0, done = 0;
    for (uint32_t b = 0; b < t->B; b++) {
        bool proof_timed_out = false;
        volatile uint64_t* status = ctx->tail_res + ctx->tail_status_off + b;
        for (uint32_t p = 0; p < t->P && !proof_timed_out; p++)
            for (uint32_t pt = 0; pt <= t->deg[p] && !proof_timed_out; pt++) {
                if (pt == 1) continue;
                const size_t e = (size_t)b * t->E + t->eoff[p] + pt;
                uint32_t limbs[8];
                for (int l = 0; l < 8 && !proof_timed_out; l++) {
                    volatile uint64_t* u = ctx->tail_res + e * 8 + l;
                    for (;;) {
                        const uint64_t v = *u;
                        const uint32_t tag = (uint32_t)(v >> 32);
                        if (tag == seq) { limbs[l] = (uint32_t)v; break; }
                        // the kernel's notices about THIS round live in the proof's status unit (sequence number | code << 32)
                        const uint64_t st = *status;
                        if ((uint32_t)st == seq && (uint32_t)(st >> 32) == kTailTimeout) { proof_timed_out = true; break; }   // gave up waiting; nothing of this round was folded
                        if ((uint32_t)st == seq && (uint32_t)(st >> 32) == kTailFailed) {
                            tail_forget(t, true);
                            FAIL(ZKSC_ERR_COMM, "resident rounds kernel: a CTA or a peer GPU went missing mid-round; the tables are undefined (reset them)");
                        }
#if defined(__x86_64__)
                        __builtin_ia32_pause();
#endif
                        if ((++spins & 0xfffff) == 0) {
                            cudaError_t q = cudaStreamQuery(ctx->stream);
                            if (q != cudaErrorNotReady && q != cudaSuccess) {
                                std::string msg = std::string("resident kernel: ") + cudaGetErrorString(q);
                                tail_forget(t, true);
                                ctx->err = msg;
                                return ZKSC_ERR_CUDA;
                            }
                            if (q == cudaSuccess && (uint32_t)(*u >> 32) != seq && (uint32_t)*status != seq) {
                                tail_forget(t, true);
                                FAIL(ZKSC_ERR_CUDA, "resident kernel ended without publishing its results");
                            }
                        }
                    }
                }
                if (!proof_timed_out && out) memcpy(out + e * 4, limbs, 32);
            }
        if (proof_timed_out) timed_out++;
        else done++;
    }
    if (timed_out == 0) return ZKSC_OK;
    if (done == 0) return kTailExpired;
    // The time-out is decided per proof (by its first CTA): some proofs of the batch folded this round and others did not, so
    // there is no single state to resume from.
    tail_forget(t, true);
    FAIL(ZKSC_ERR_COMM, "resident rounds kernel: the host's challenge reached only part of the batch in time; the tables are undefined (reset them)");
}

// bookkeeping after a resident round's result has been collected: its fold has been applied
static void tail_round_done(zksc_tables* t) {
    zksc_ctx* ctx = t->ctx;
    const unsigned int idx = t->tail_done++;
    t->last_partial = (ctx->n_ranks > 1 && t->where != 2 && (t->tail_gather_round == kNoGather || idx <= t->tail_gather_round));
    if (t->where == 0) t->where = 1;
    t->cur_n /= 2;
    if (t->tail_gather_round != kNoGather && idx == t->tail_gather_round + 1) {
        // that round pulled every rank's shard into `tail` and folded it: replicated from here on
        t->where = 2;
        t->cur_n = t->tail_gather_local * ctx->n_ranks / 2;
    }
    t->pending = false;
    t->tail_posted = false;
    t->tail_seen = std::chrono::steady_clock::now();
    if (--t->tail_left == 0) {
        t->tail_running = false;          // the kernel leaves by itself after its last round
        if (ctx->active_tail == t) ctx->active_tail = nullptr;
    }
}

// Stop a resident kernel (anything else that wants the stream, the tables or a device-wide call must do this first).  A posted
// round is completed and accounted for; its evaluations are dropped.
static int tail_stop(zksc_tables* t) {
    if (!t || !t->tail_running) return ZKSC_OK;
    zksc_ctx* ctx = t->ctx;
    CK(cudaSetDevice(ctx->device));
    if (t->tail_posted) {
        const int rc = tail_wait(t, t->tail_cur, nullptr);
        if (rc == kTailExpired) return tail_expired(t);
        if (rc != ZKSC_OK) return rc;          // tail_wait has cleaned up
        tail_round_done(t);
        t->claim_valid = false;
        t->last_evals_valid = false;
    }
    if (t->tail_running) {
        for (uint32_t b = 0; b < t->B; b++) ctx->tail_mail[(size_t)b * kMailUnits] = (uint64_t)kTailAbort << 32;
        t->tail_running = false;
        t->tail_left = 0;
    }
    if (ctx->active_tail == t) ctx->active_tail = nullptr;
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t b = 0; b < t->B && b < ctx->tail_proofs_cap; b++)       // the kernel is gone: no "leave" mark stays behind
        if ((uint32_t)(ctx->tail_mail[(size_t)b * kMailUnits] >> 32) == kTailAbort) ctx->tail_mail[(size_t)b * kMailUnits] = 0;
    return ZKSC_OK;
}
static int quiesce(zksc_ctx* ctx) { return ctx && ctx->active_tail ? tail_stop(ctx->active_tail) : ZKSC_OK; }

// which instantiation of the resident kernel serves this handle: its products' common degree, 0 when they differ
static int tail_dsel(const zksc_tables* t) {
    for (uint32_t p = 1; p < t->P; p++)
        if (t->deg[p] != t->deg[0]) return 0;
    return (int)t->deg[0];
}
// CTAs per (proof, product) group: the whole grid must be co-resident (cooperative launch), every group gets the same share
static unsigned int tail_group_ctas(const zksc_tables* t) {
    const zksc_ctx* ctx = t->ctx;
    const int dsel = tail_dsel(t);
    if (dsel > kResMaxDegree) return 0;
    const size_t groups = (size_t)t->B * t->P;
    const size_t cap = (size_t)ctx->sms * (size_t)std::max(ctx->res_occ[dsel], 0);
    return groups ? (unsigned int)(cap / groups) : 0;
}
// 32 x 32 limb products per pair of a fused fold + evaluate round
static unsigned long long round_products_per_pair(unsigned long long d) {
    const unsigned long long per_point = d == 1 ? 0 : (d <= 3 ? (d - 2) * 120 + 64 : (d - 1) * 120);
    return 2 * d * 76 + d * per_point;
}
static bool tail_eligible(const zksc_tables* t, unsigned long long half, bool sharded) {
    const zksc_ctx* ctx = t->ctx;
    if (!ctx->tail_enabled || tail_group_ctas(t) == 0) return false;
    if (half >= (1ull << 29) || geo_of(t, 0).tab_stride >= (1ull << 31)) return false;     // the kernel indexes with 32 bits
    if (sharded) {
        if (!ctx->p2p || (size_t)t->B * t->E > kXchCap) return false;
        if ((size_t)t->B * t->Dtot * 2 > kXchStageElems) return false;                      // the gather stage must hold at least one pair per table
    }
    // the resident kernel reads the fold table from shared memory instead of the constant bank and keeps fewer loads in flight: a few
    // per cent slower per pair than the ordinary launch, so it takes over where a round's fixed costs outweigh that
    unsigned long long work = 0;
    for (uint32_t p = 0; p < t->P; p++) {
        if (t->deg[p] > (uint32_t)kResMaxDegree) return false;
        work += half * round_products_per_pair(t->deg[p]);
    }
    return work * t->B <= ctx->tail_work;
}

// Launch the resident kernel for all remaining rounds; the pending challenge is its first mailbox message.
// prelaunch: the kernel is queued BEHIND the ordinary launch of the round before its first one, while that round still runs -- the
// cooperative launch (~12 us of API time and start-up, against ~5 for an ordinary one) disappears behind it; the first challenge does not
// exist yet, the kernel waits for it in the mailbox like for every later one (zksc_bind posts it).
static int tail_start(zksc_tables* t, unsigned long long half, bool sharded, bool prelaunch = false) {
    zksc_ctx* ctx = t->ctx;
    TRY(quiesce(ctx));
    const unsigned int cpg = tail_group_ctas(t);
    const size_t groups = (size_t)t->B * t->P;
    TRY(tail_ensure(ctx, t->B, (size_t)t->B * t->E * 8, groups, groups * kResMaxDegree * cpg));
    const unsigned int G = sharded ? (unsigned int)ctx->n_ranks : 1u;
    // rounds: the table before this kernel's first fold has 4 * half * G entries in total = 2^m; m - 1 rounds are left
    unsigned int n_rounds = 0;
    for (unsigned long long h = half * G; h >= 1; h >>= 1) n_rounds++;
    unsigned int gather_round = kNoGather;
    unsigned long long gather_local = 0;
    if (sharded) {
        // gather when the table (after the fold) is down to `want` entries in total; at least one pair per rank must be left, and
        // every rank's shard of every table must fit its stage
        unsigned long long want = ctx->gather_entries;
        while (want > 2ull * G && (size_t)t->B * t->Dtot * (want / G) > kXchStageElems) want /= 2;
        unsigned long long total = 2 * half * G;      // after the first round's fold
        gather_round = 0;
        while (total > want) { total /= 2; gather_round++; }
        gather_local = total / G;
    }
    if (ctx->tail_seq > 0xf0000000u) {                    // stay clear of the reserved sequence numbers
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->tail_seq = 0;
        memset((void*)ctx->tail_mail, 0, ctx->tail_proofs_cap * kMailUnits * 8);
        memset((void*)ctx->tail_res, 0, (ctx->tail_units_cap + ctx->tail_proofs_cap) * 8);
        CK(cudaMemsetAsync(ctx->tail_relay, 0, 2 * ctx->tail_proofs_cap * (kMailUnits + 1) * sizeof(unsigned long long), ctx->stream));
        if (ctx->xch_local) CK(cudaMemsetAsync(ctx->xch_local + kXchUnitsOffset(ctx->n_ranks), 0, kXchUnitsBytes(ctx->n_ranks), ctx->stream));
    }
    const unsigned int seq0 = ctx->tail_seq + 1;
    ctx->tail_seq += n_rounds;
    Geo gi = geo_of(t, t->where);
    Geo go = geo_of(t, t->where == 0 ? 1 : t->where);
    ResArgs a;
    memset(&a, 0, sizeof(a));
    a.in = gi.base; a.out = go.base;
    a.in_tab_stride = gi.tab_stride; a.in_proof_stride = gi.proof_stride;
    a.out_tab_stride = go.tab_stride; a.out_proof_stride = go.proof_stride;
    a.half = half; a.n_rounds = n_rounds; a.seq0 = seq0;
    a.n_proofs = t->B; a.n_products = t->P; a.n_evals = t->E; a.n_tables = t->Dtot;
    for (uint32_t p = 0; p < t->P; p++) { a.deg[p] = t->deg[p]; a.koff[p] = t->koff[p]; a.eoff[p] = t->eoff[p]; }
    a.cpg = cpg;
    a.mail = ctx->tail_mail_dev; a.results = ctx->tail_res_dev; a.status = ctx->tail_res_dev + ctx->tail_status_off;
    a.relay = ctx->tail_relay; a.relay_tags = ctx->tail_relay + 2 * ctx->tail_proofs_cap * kMailUnits;
    a.partials = ctx->tail_partials; a.counters = ctx->tail_counters;
    a.work = ctx->tail_dynamic ? ctx->tail_counters + kTailWorkBase(ctx->tail_groups_cap) : nullptr;
    a.n_ranks = 1; a.rank = 0; a.xch_cap = kXchCap;
    a.gather_round = kNoGather;
    if (sharded) {
        a.n_ranks = ctx->n_ranks; a.rank = ctx->rank;
        for (int g = 0; g < ctx->n_ranks; g++) {
            a.peer_units[g] = (unsigned long long*)((unsigned char*)ctx->xch_peer[g] + kXchUnitsOffset(ctx->n_ranks));   // unused with host_reduce
            a.peer_stage[g] = (const Fr*)((unsigned char*)ctx->xch_peer[g] + kXchStageOffset(ctx->n_ranks));
        }
        a.gather_round = gather_round; a.gather_local = gather_local;
        a.host_reduce = ctx->host_reduce ? 1u : 0u;
        Geo gt = geo_of(t, 2);
        a.tail = gt.base; a.tail_tab_stride = gt.tab_stride; a.tail_proof_stride = gt.proof_stride;
    }
    a.relay_cap = (unsigned int)ctx->tail_proofs_cap;
    if (prelaunch) {
        // nothing is posted before this launch, so whatever an earlier session left in the mailbox is what the kernel reads first: a
        // "leave" mark of a session that was stopped half-way must not be there (sequence number 0 is never used)
        for (uint32_t b = 0; b < t->B; b++) ctx->tail_mail[(size_t)b * kMailUnits] = 0;
        t->tail_cur = seq0 - 1;
        t->tail_posted = false;
        t->tail_seen = std::chrono::steady_clock::now();
    } else {
        tail_post(t, seq0);
    }
    cudaError_t le = zksc_launch_resident(tail_dsel(t), (unsigned int)(groups * cpg), ctx->stream, a);
    if (le != cudaSuccess) {
        // cannot be made resident (another context holds the SMs, MPS limits, ...): ordinary launches from now on
        cudaGetLastError();
        t->tail_posted = false;
        ctx->tail_enabled = false;
        return kTailExpired;
    }
    ctx->launches++;
    t->tail_running = true;
    t->tail_left = n_rounds;
    t->tail_done = 0;
    t->tail_gather_round = gather_round;
    t->tail_gather_local = gather_local;
    ctx->active_tail = t;
    return ZKSC_OK;
}

// Apply the pending challenge with the stand-alone fold kernel (no evaluation).
static int flush_pending(zksc_tables* t) {
    zksc_ctx* ctx = t->ctx;
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    if (!t->pending) return ZKSC_OK;
    if (t->cur_n < 2) FAIL(ZKSC_ERR_STATE, "no variable left to bind");
    Geo gi = geo_of(t, t->where);
    int to = (t->where == 0) ? 1 : t->where;
    Geo go = geo_of(t, to);
    for (uint32_t b0 = 0; b0 < t->B; b0 += kMaxBatch) {
        uint32_t nb = t->B - b0 < (uint32_t)kMaxBatch ? t->B - b0 : kMaxBatch;
        FoldArgs a;
        a.in = gi.base + (size_t)b0 * gi.proof_stride;
        a.out = go.base + (size_t)b0 * go.proof_stride;
        a.in_tab_stride = gi.tab_stride; a.in_proof_stride = gi.proof_stride;
        a.out_tab_stride = go.tab_stride; a.out_proof_stride = go.proof_stride;
        a.n_out = t->cur_n / 2; a.s = t->cur_n / 2; a.n_tabs = t->Dtot;
        for (uint32_t b = 0; b < nb; b++) a.chal[b] = t->pending_chal[b0 + b];
        dim3 grid(grid_for(ctx, a.n_out, 256, 8), nb);
        fold_kernel<<<grid, 256, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    t->where = to;
    t->cur_n /= 2;
    t->pending = false;
    return ZKSC_OK;
}

#if ZKSC_HAVE_NCCL_H
#define NCCLCK(call)                                                                       \
    do {                                                                                   \
        ncclResult_t r_ = (call);                                                          \
        if (r_ != ncclSuccess) FAIL(ZKSC_ERR_COMM, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)
#endif

__global__ void tail_scatter_kernel(const Fr* gathered, Fr* tail, unsigned long long tail_stride, unsigned int n_tabs_total, unsigned int G) {
    // gathered: [rank][tab] -> tail: [tab][rank] (tables tail_stride elements apart)
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tabs_total * G) {
        unsigned int tab = i / G, g = i % G;
        st256(tail + (size_t)tab * tail_stride + g, ld256(gathered + (size_t)g * n_tabs_total + tab));
    }
}
__global__ void tail_collect_kernel(const Fr* base, unsigned long long tab_stride, Fr* out, unsigned int n_tabs_total) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tabs_total) st256(out + i, ld256(base + (size_t)i * tab_stride));
}

// Sharded contexts: once every rank is down to one entry per table, gather the G entries of every
// table on every rank (entry index == rank) and continue replicated.
static int gather_tail(zksc_tables* t) {
    zksc_ctx* ctx = t->ctx;
#if ZKSC_HAVE_NCCL_H
    TRY(flush_pending(t));
    if (t->cur_n != 1) FAIL(ZKSC_ERR_STATE, "gather_tail: local tables not exhausted");
    const unsigned int nt = t->B * t->Dtot;
    TRY(ensure_results(ctx, nt));
    Geo g = geo_of(t, t->where);
    tail_collect_kernel<<<(nt + 127) / 128, 128, 0, ctx->stream>>>(g.base, g.tab_stride, ctx->results_send, nt);
    ctx->launches++;
    CK(cudaGetLastError());
    NCCLCK(g_nccl.AllGather(ctx->results_send, ctx->results_dev, (size_t)nt * sizeof(Fr), ncclUint8, ctx->comm, ctx->stream));
    tail_scatter_kernel<<<(nt * ctx->n_ranks + 127) / 128, 128, 0, ctx->stream>>>(ctx->results_dev, t->tail, t->tail_stride, nt, ctx->n_ranks);
    ctx->launches++;
    CK(cudaGetLastError());
    t->where = 2;
    t->cur_n = ctx->n_ranks;
    return ZKSC_OK;
#else
    FAIL(ZKSC_ERR_COMM, "built without nccl.h");
#endif
}

// Wait for the round's completion word (written by the last block of the round's last launch, after its
// results, with a system-scope fence in between).  Polling a pinned word costs ~1 us after the kernel's
// store; cudaMemcpyAsync + cudaStreamSynchronize cost ~15 us per round.  The stream is queried now and then so
// that a failed launch surfaces as an error instead of an endless spin.
static int wait_flag(zksc_ctx* ctx, unsigned int seq) {
    for (unsigned long long spins = 1;; spins++) {
        const unsigned int f = *ctx->flag_host;
        if (f == seq) break;
        if (f == 0xffffffffu) { *ctx->flag_host = 0; FAIL(ZKSC_ERR_COMM, "timed out waiting for a peer GPU's partial evaluations"); }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0xffff) == 0) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaErrorNotReady) continue;
            if (q != cudaSuccess) { ctx->err = std::string("round kernel: ") + cudaGetErrorString(q); return ZKSC_ERR_CUDA; }
            if (*ctx->flag_host == seq) break;
            FAIL(ZKSC_ERR_CUDA, "round kernel finished without publishing its results");
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    return ZKSC_OK;
}

// fill in the per-proof challenge data and launch the round kernel of degree D
template <int NB>
static int launch_round(const zksc_tables* t, const RoundBase& base, int D, int variant, bool staged, dim3 grid, bool fold, uint32_t b0, uint32_t nb) {
    static thread_local RoundArgsT<NB> a;   // 18 KiB for NB = kMaxBatch: kept off the stack
    static_cast<RoundBase&>(a) = base;
    if (fold) for (uint32_t b = 0; b < nb; b++) { a.chal[b] = t->pending_chal[b0 + b]; a.tab[b] = t->pending_tab[b0 + b]; }
    cudaStream_t s = t->ctx->stream;
    switch (D) {
        case 1: zksc_launch_round_1(variant, staged, grid, s, a); break;
        case 2: zksc_launch_round_2(variant, staged, grid, s, a); break;
        case 3: zksc_launch_round_3(variant, staged, grid, s, a); break;
        case 4: zksc_launch_round_4(variant, staged, grid, s, a); break;
        case 5: zksc_launch_round_5(variant, staged, grid, s, a); break;
        case 6: zksc_launch_round_6(variant, staged, grid, s, a); break;
        case 7: zksc_launch_round_7(variant, staged, grid, s, a); break;
        case 8: zksc_launch_round_8(variant, staged, grid, s, a); break;
        default: return ZKSC_ERR_UNSUPPORTED;
    }
    return ZKSC_OK;
}

// after the device part of a round: fill in point 1 from the claim, remember the evaluations for the next claim
static void finish_round(zksc_tables* t, uint64_t* out, bool skip1, bool full, uint32_t npts_cap) {
    if (t->raw) {          // a child of a multi-GPU handle returns its shard's sums as they are; the parent finishes the totals
        t->claim_valid = false;
        t->last_evals_valid = full;
        return;
    }
    if (skip1) {
        // h_p(1) = claim_p - h_p(0)   (what the verifier checks; exact in the field)
        for (uint32_t b = 0; b < t->B; b++)
            for (uint32_t p = 0; p < t->P; p++) {
                uint64_t* e = out + ((size_t)b * t->E + t->eoff[p]) * 4;
                store_h(e + 4, host::sub(t->claim[(size_t)b * t->P + p], load_h(e)));
            }
    }
    // degree-2 and degree-3 products: the kernels deliver h(inf) (and h(-1)) instead of h(2) (and h(3)); see
    // host::round_slots_to_evals.  Every caller asks for all degree + 1 values (npts_cap = ZKSC_MAX_DEGREE + 1).
    for (uint32_t p = 0; p < t->P; p++) {
        const uint32_t d = t->deg[p];
        if ((d != 2 && d != 3) || npts_cap <= d) continue;
        for (uint32_t b = 0; b < t->B; b++) host::round_slots_to_evals(d, out + ((size_t)b * t->E + t->eoff[p]) * 4);
    }
    t->claim_valid = false;
    if (full) {
        t->last_evals.assign(out, out + (size_t)t->B * t->E * 4);
        t->last_evals_valid = true;
    }
}

// One round: evaluations of every product of every proof at 0..npts_cap-1 (capped by degree+1).
static int multi_round_evals(zksc_tables* t, uint64_t* out, uint32_t npts_cap);
static int round_evals_impl(zksc_tables* t, uint64_t* out, uint32_t npts_cap) {
    zksc_ctx* ctx = t->ctx;
    if (!t->kids.empty()) return multi_round_evals(t, out, npts_cap);
    CK(cudaSetDevice(ctx->device));
    TRY(copy_finish(t));
    if (t->vars_left == 0) FAIL(ZKSC_ERR_STATE, "all variables are bound");
    if (t->r0_valid && t->vars_left == t->n_vars && !t->pending && npts_cap > ZKSC_MAX_DEGREE) {
        memcpy(out, t->r0_cache.data(), t->r0_cache.size() * sizeof(uint64_t));
        t->r0_valid = false;
        t->last_evals = t->r0_cache;
        t->last_evals_valid = true;
        return ZKSC_OK;
    }
    const auto prof_t0 = std::chrono::steady_clock::now();
    if (ctx->active_tail && ctx->active_tail != t) TRY(quiesce(ctx));
    if (t->tail_running && (!t->tail_posted || npts_cap <= ZKSC_MAX_DEGREE)) TRY(tail_stop(t));   // not the prover's call pattern
    // the rounds of a running resident kernel: collect the posted round's evaluations (the kernel knows its own table geometry --
    // on a sharded context the local tables may be down to one pair while the gathered table it works on next is larger)
    auto resident_round = [&]() -> int {
        const auto w0 = std::chrono::steady_clock::now();
        ctx->prof_launch = std::chrono::duration<double, std::micro>(w0 - prof_t0).count();
        int rc = tail_wait(t, t->tail_cur, out);
        if (rc == ZKSC_OK) {
            ctx->prof_wait = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - w0).count();
            tail_round_done(t);
            finish_round(t, out, true, true, npts_cap);
            return ZKSC_OK;
        }
        if (rc != kTailExpired) return rc;
        TRY(tail_expired(t));                         // the kernel gave up waiting for the host: this round goes through an ordinary launch
        return round_evals_impl(t, out, npts_cap);    // (the resident kernel is off for this context now)
    };
    if (t->tail_running) return resident_round();
    const bool sharded_phase = (ctx->n_ranks > 1 && t->where != 2);
    if (sharded_phase) {
        uint64_t n_after = t->pending ? t->cur_n / 2 : t->cur_n;
        if (n_after == 1) {
            if (ctx->host_reduce) FAIL(ZKSC_ERR_STATE, "multi-GPU child asked to gather (the parent does that)");
            TRY(gather_tail(t));
        }
    }
    const bool reduce_ranks = (ctx->n_ranks > 1 && t->where != 2);
    uint64_t n_eval = t->pending ? t->cur_n / 2 : t->cur_n;  // table size of the round being evaluated
    if (n_eval < 2) FAIL(ZKSC_ERR_STATE, "no variable left to evaluate");
    const bool fold = t->pending;
    const bool full = npts_cap > ZKSC_MAX_DEGREE;
    const bool skip1 = fold && full && t->claim_valid;      // point 1 comes from the claim
    const int variant = skip1 ? 2 : (fold ? 1 : 0);
    Geo gi = geo_of(t, t->where);
    int to = fold ? ((t->where == 0) ? 1 : t->where) : t->where;
    Geo go = geo_of(t, to);
    const unsigned long long half = n_eval / 2;
    const size_t n_res = (size_t)t->B * t->E;
    if (skip1 && tail_eligible(t, half, reduce_ranks)) {
        // latency-bound rounds: the resident kernel runs this round and all later ones
        const int rc = tail_start(t, half, reduce_ranks);
        if (rc == ZKSC_OK) return resident_round();
        if (rc != kTailExpired) return rc;
        return round_evals_impl(t, out, npts_cap);    // could not be made resident: off for this context, ordinary launches from here
    }
    t->last_partial = reduce_ranks;
    const bool peer = reduce_ranks && ctx->p2p && !ctx->host_reduce && n_res <= kXchCap;   // exchange + sum inside the round kernel
    const bool mapped = (ctx->mapped_results && !reduce_ranks) || peer || ctx->host_reduce;
    Fr* res = (reduce_ranks && !ctx->host_reduce) ? ctx->results_send : (mapped ? ctx->results_host_dev : ctx->results_dev);
    const unsigned int seq = ++ctx->flag_seq;
    if (peer) ctx->xch_seq++;

    // products of one degree share a launch (blockIdx.z = product): GKR's two degree-2 products cost one launch per round
    bool same_degree = t->P > 1 && ctx->fuse_products;
    for (uint32_t p = 1; p < t->P; p++) same_degree = same_degree && t->deg[p] == t->deg[0];
    const uint32_t p_step = same_degree ? t->P : 1;
    for (uint32_t b0 = 0; b0 < t->B; b0 += kMaxBatch) {
        uint32_t nb = t->B - b0 < (uint32_t)kMaxBatch ? t->B - b0 : kMaxBatch;
        for (uint32_t p = 0; p < t->P; p += p_step) {
            const int D = t->deg[p];
            // the staged kernel works on whole 32-pair warp tiles and pays off once HBM latency matters
            const bool staged = ctx->staged && (variant == 0 || ctx->staged_fold) && ctx->occ[D][3 + variant] > 0 && half % kTilePairs == 0 && half >= 4096;
            int gx = staged ? grid_for(ctx, half / kTilePairs * 32, kThreads, ctx->occ[D][3 + variant]) : grid_for(ctx, half, kThreads, ctx->occ[D][variant]);
            // Chunks from a counter (kernels.cuh RoundBase::dynamic) need every CTA of a group resident at once -- the first ones to
            // start would take all of the group's work -- so it is for launches of a few groups, the grid cut to what fits; a batch of
            // many proofs is balanced by the CTA scheduler itself (37888 short-lived CTAs for 64 proofs), with a fixed split.
            const unsigned int n_groups = nb * p_step;
            const bool dynamic = ctx->round_dynamic && n_groups <= ctx->dyn_max_groups && D >= 2;     // degree 1 is bound by HBM alone: nothing to balance
            if (dynamic) gx = std::max(1, std::min(gx, ctx->sms * ctx->occ[D][staged ? 3 + variant : variant] / (int)n_groups));
            TRY(ensure_partials(ctx, (size_t)nb * gx * (D + 1) * p_step));
            RoundBase base;
            base.in = gi.base + (size_t)b0 * gi.proof_stride + (size_t)t->koff[p] * gi.tab_stride;
            base.out = go.base + (size_t)b0 * go.proof_stride + (size_t)t->koff[p] * go.tab_stride;
            base.in_tab_stride = gi.tab_stride; base.in_proof_stride = gi.proof_stride;
            base.out_tab_stride = go.tab_stride; base.out_proof_stride = go.proof_stride;
            base.in_prod_stride = (unsigned long long)D * gi.tab_stride; base.out_prod_stride = (unsigned long long)D * go.tab_stride;
            base.res_prod_stride = (unsigned int)(D + 1);
            base.half = half;
            base.partials = ctx->partials; base.counters = ctx->counters;
            base.dynamic = dynamic ? 1u : 0u;
            base.result = res + (size_t)b0 * t->E + t->eoff[p];
            base.res_stride = t->E;
            base.npts = (uint32_t)(D + 1) < npts_cap ? (D + 1) : npts_cap;
            const bool last_launch = (b0 + nb == t->B) && (p + p_step == t->P);
            base.flag = (mapped && last_launch) ? ctx->flag_dev : nullptr;
            base.flag_value = seq;
            base.xch.n_ranks = 0;
            if (peer && last_launch) {
                XchArgs& x = base.xch;
                for (int g = 0; g < ctx->n_ranks; g++) {
                    x.peer_tags[g] = (unsigned int*)ctx->xch_peer[g];
                    x.peer_data[g] = (Fr*)((unsigned char*)ctx->xch_peer[g] + kXchTagBytes);
                }
                x.send = ctx->results_send; x.out = ctx->results_host_dev;
                x.n_ranks = ctx->n_ranks; x.rank = ctx->rank; x.seq = ctx->xch_seq; x.n_elems = (unsigned int)n_res; x.cap = kXchCap;
            }
            dim3 grid(gx, nb, p_step);
            TRY(timing_open(ctx, D, fold, half, (unsigned long long)nb * p_step));
            int rc = (nb == 1) ? launch_round<1>(t, base, D, variant, staged, grid, fold, b0, nb) : launch_round<kMaxBatch>(t, base, D, variant, staged, grid, fold, b0, nb);
            if (rc != ZKSC_OK) FAIL(rc, "degree");
            TRY(timing_close(ctx));
            ctx->launches++;
            CK(cudaGetLastError());
        }
    }
    if (fold) { t->where = to; t->cur_n /= 2; t->pending = false; }

    if (mapped) {
        // The next round will be the resident kernel's first: queue it now, behind this round's launch (see tail_start).  Only where nothing
        // else of this round is queued after this point (host-mapped results), on unsharded contexts, in the prover's call pattern.
        if (ctx->tail_prelaunch && t->in_prove && full && ctx->n_ranks == 1 && !t->tail_running && t->cur_n >= 4 && tail_eligible(t, t->cur_n / 4, false)) {
            const int rc = tail_start(t, t->cur_n / 4, false, true);
            if (rc != ZKSC_OK && rc != kTailExpired) return rc;      // (expired: cannot be made resident; the context is on ordinary launches now)
        }
        const auto w0 = std::chrono::steady_clock::now();
        ctx->prof_launch = std::chrono::duration<double, std::micro>(w0 - prof_t0).count();
        TRY(wait_flag(ctx, seq));
        ctx->prof_wait = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - w0).count();
        memcpy(out, ctx->results_host, n_res * sizeof(Fr));
    } else if (!reduce_ranks) {
        CK(cudaMemcpyAsync(ctx->results_host, ctx->results_dev, n_res * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        memcpy(out, ctx->results_host, n_res * sizeof(Fr));
    } else {
#if ZKSC_HAVE_NCCL_H
        // the only per-round exchange: every rank's partial evaluations (n_res elements of 32 bytes).
        // NCCL has no modular sum, so gather and add mod r on the host, identically on every rank.
        NCCLCK(g_nccl.AllGather(ctx->results_send, ctx->results_dev, n_res * sizeof(Fr), ncclUint8, ctx->comm, ctx->stream));
        CK(cudaMemcpyAsync(ctx->results_host, ctx->results_dev, n_res * ctx->n_ranks * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < n_res; i++) {
            FrH acc = to_host(ctx->results_host[i]);
            for (int g = 1; g < ctx->n_ranks; g++) acc = host::add(acc, to_host(ctx->results_host[(size_t)g * n_res + i]));
            store_h(out + 4 * i, acc);
        }
#endif
    }
    finish_round(t, out, skip1, full, npts_cap);
    return ZKSC_OK;
}

extern "C" int zksc_round_evals(zksc_tables* t, uint64_t* out) {
    if (!t || !out) return ZKSC_ERR_STATE;
    return round_evals_impl(t, out, ZKSC_MAX_DEGREE + 1);
}

// fn(b) for every proof of the handle: on the worker pool while a batched zksc_prove is running, inline otherwise
static void for_each_proof(zksc_ctx* ctx, uint32_t B, const std::function<void(uint32_t)>& fn) {
    if (ctx->pool && ctx->pool_active && B >= kPoolMinProofs) ctx->pool->parallel_for(B, fn);
    else for (uint32_t b = 0; b < B; b++) fn(b);
}

static int multi_bind(zksc_tables* t, const uint64_t* challenges);
extern "C" int zksc_bind(zksc_tables* t, const uint64_t* challenges) {
    if (!t || !challenges) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (!t->kids.empty()) return multi_bind(t, challenges);
    CK(cudaSetDevice(ctx->device));
    if (t->vars_left == 0) FAIL(ZKSC_ERR_STATE, "all variables are bound");
    if (ctx->active_tail && ctx->active_tail != t) TRY(quiesce(ctx));
    if (t->tail_running && (t->tail_posted || !t->last_evals_valid)) TRY(tail_stop(t));   // bind without the round's evaluations
    if (t->pending) {
        if (ctx->n_ranks > 1 && !ctx->host_reduce && t->where != 2 && t->cur_n / 2 == 1) TRY(gather_tail(t));
        else TRY(flush_pending(t));
    }
    if (ctx->n_ranks > 1 && t->where != 2 && t->cur_n == 1) {
        if (ctx->host_reduce) FAIL(ZKSC_ERR_STATE, "multi-GPU child asked to gather (the parent does that)");
        TRY(gather_tail(t));
    }
    t->pending_chal.resize(t->B);
    t->pending_tab.resize(t->B);
    // per proof: the fold table of its challenge and, for the next round, the claim = this round's polynomial of every product
    // at the challenge
    const bool claims = t->raw ? t->raw_claim_next : t->last_evals_valid;
    if (claims && !t->raw) t->claim.resize((size_t)t->B * t->P);
    bool post = false;
    if (t->tail_running) {
        // The kernel's CTAs give up waiting for a challenge after kTailTimeoutNs, each by its own clock.  A challenge posted close to
        // that deadline could reach some and not others, so the host never posts later than half of it after it saw the results:
        // past that, the kernel is told to leave (nothing of the round has been folded) and ordinary launches carry on.
        const double waited_ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t->tail_seen).count();
        if (waited_ns > 0.5 * (double)kTailTimeoutNs) {
            // (also what happens when launches are synchronous -- a profiler -- and the kernel was queued ahead of its first round: it has
            // come and gone while the launch call blocked.  Either way this host cannot feed a resident kernel: ordinary launches for good.)
            TRY(tail_stop(t));
            ctx->tail_enabled = false;
        } else post = true;
    }
    const unsigned int post_seq = t->tail_cur + 1;
    for_each_proof(ctx, t->B, [&](uint32_t b) {
        memcpy(t->pending_chal[b].l, challenges + 4 * b, 32);
        host::fold_table(load_h(challenges + 4 * b), t->pending_tab[b].w);
        if (post) {
            // the resident kernel folds with it and evaluates the next round: posted as soon as the table exists, before the claims
            // below (which only the NEXT collection needs) -- this store is on every round's critical path
            const FoldTab& w = t->pending_tab[b];
            volatile uint64_t* m = ctx->tail_mail + (size_t)b * kMailUnits;
            for (int i = 0; i < 8; i++)
                for (int j = 0; j < 8; j++) m[foldtabs_index(i, j)] = (uint64_t)w.w[i][j] | ((uint64_t)post_seq << 32);
        }
        if (!claims || t->raw) return;
        std::vector<FrH> ys;
        for (uint32_t p = 0; p < t->P; p++) {
            ys.clear();
            for (uint32_t i = 0; i <= t->deg[p]; i++) ys.push_back(load_h(&t->last_evals[((size_t)b * t->E + t->eoff[p] + i) * 4]));
            t->claim[(size_t)b * t->P + p] = host::SparseUnivariatePolynomial::evaluate_evals_at(ys, load_h(challenges + 4 * b));
        }
    });
    t->pending = true;
    t->vars_left--;
    t->claim_valid = claims;
    t->last_evals_valid = false;
    if (post) { t->tail_cur = post_seq; t->tail_posted = true; }
    return ZKSC_OK;
}

extern "C" int zksc_residual(zksc_tables* t, uint64_t* out) {
    if (!t || !out) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (!t->kids.empty()) return multi_residual(t, out);
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    if (ctx->n_ranks > 1 && t->where != 2) {
        uint64_t n_after = t->pending ? t->cur_n / 2 : t->cur_n;
        if (n_after != 1) FAIL(ZKSC_ERR_UNSUPPORTED, "sharded residual is only available once each rank holds one entry per table");
        TRY(gather_tail(t));
    }
    TRY(flush_pending(t));
    Geo g = geo_of(t, t->where);
    const size_t nt = (size_t)t->B * t->Dtot;
    for (size_t i = 0; i < nt; i++)
        CK(cudaMemcpyAsync(out + 4 * i * t->cur_n, g.base + i * g.tab_stride, t->cur_n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}


// ------------------------------------------------------------------------------------------------
// single-process multi-GPU: one host thread, G devices (the reference's caller, gkr/src/protocol.rs:85, is one process)
// ------------------------------------------------------------------------------------------------
// fn(g) on every child, concurrently (one pool thread per device) inside a GpuPoolSession, one after the other otherwise
static int for_each_kid(zksc_ctx* ctx, const std::function<int(uint32_t)>& fn) {
    const uint32_t G = (uint32_t)ctx->kids.size();
    std::vector<int> rc(G, ZKSC_OK);
    const std::function<void(uint32_t)> body = [&](uint32_t g) { rc[g] = fn(g); };
    if (ctx->gpu_pool && ctx->gpu_pool_active) ctx->gpu_pool->parallel_for(G, body);
    else for (uint32_t g = 0; g < G; g++) body(g);
    for (uint32_t g = 0; g < G; g++)
        if (rc[g] != ZKSC_OK) { ctx->err = "device " + std::to_string(ctx->kids[g]->device) + ": " + ctx->kids[g]->err; return rc[g]; }
    return ZKSC_OK;
}

extern "C" int zksc_ctx_create_multi(const int* devices, int n_devices, zksc_ctx** out) {
    if (!out) return ZKSC_ERR_SHAPE;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > kMaxRanks || (n_devices & (n_devices - 1))) {
        g_create_error = "zksc_ctx_create_multi: the device count must be a power of two, at most 8";
        return ZKSC_ERR_SHAPE;
    }
    for (int a = 0; a < n_devices; a++)
        for (int b = 0; b < a; b++)
            if (devices[a] == devices[b]) { g_create_error = "zksc_ctx_create_multi: the devices must be distinct"; return ZKSC_ERR_SHAPE; }
    if (n_devices == 1) return zksc_ctx_create(devices[0], out);
    zksc_ctx* ctx = new zksc_ctx();
    auto fail = [&](int rc, const std::string& why) {
        g_create_error = why;
        zksc_ctx_destroy(ctx);
        return rc;
    };
    const int G = n_devices;
    for (int g = 0; g < G; g++) {
        zksc_ctx* k = nullptr;
        const int rc = zksc_ctx_create(devices[g], &k);
        if (rc != ZKSC_OK) {
            if (ctx->kids.empty()) { delete ctx; return rc; }      // g_create_error is set
            return fail(rc, g_create_error);
        }
        k->parent = ctx; k->rank = g; k->n_ranks = G; k->host_reduce = true;
        if (!getenv("ZKSC_TAIL_WORK")) k->tail_work = kTailWorkSharded;
        ctx->kids.push_back(k);
    }
    ctx->device = devices[0];
    ctx->sms = ctx->kids[0]->sms;
    {
        unsigned long long ge = kGatherDefault;
        const char* e_ = getenv("ZKSC_GATHER_ENTRIES");
        if (e_ && atoll(e_) > 0) ge = (unsigned long long)atoll(e_);
        while (ge & (ge - 1)) ge &= ge - 1;
        if (ge < 2ull * G) ge = 2ull * G;
        ctx->gather_entries = ge;
        for (zksc_ctx* k : ctx->kids) k->gather_entries = ge;
    }
    // every device reads every other device's gather stage (and device 0's upload staging) directly: peer access over NVLink
    const size_t bytes = kXchStageOffset(G) + kXchStageElems * sizeof(Fr);
    for (int a = 0; a < G; a++) {
        cudaError_t e = cudaSetDevice(devices[a]);
        for (int b = 0; b < G && e == cudaSuccess; b++) {
            if (a == b) continue;
            int can = 0;
            e = cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (e == cudaSuccess && !can) return fail(ZKSC_ERR_UNSUPPORTED, "zksc_ctx_create_multi: the devices cannot access each other's memory (no NVLink / PCIe peer path)");
            if (e == cudaSuccess) {
                e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            }
        }
        if (e == cudaSuccess) e = cudaMalloc(&ctx->kids[a]->xch_local, bytes);
        if (e == cudaSuccess) e = cudaMemset(ctx->kids[a]->xch_local, 0, bytes);
        if (e != cudaSuccess) return fail(ZKSC_ERR_CUDA, std::string("zksc_ctx_create_multi: ") + cudaGetErrorString(e));
    }
    for (int a = 0; a < G; a++) {
        for (int b = 0; b < G; b++) ctx->kids[a]->xch_peer[b] = ctx->kids[b]->xch_local;
        ctx->kids[a]->p2p = true;
    }
    ctx->gpu_pool = new HostPool(G - 1);
    *out = ctx;
    return ZKSC_OK;
}

extern "C" int zksc_ctx_devices(const zksc_ctx* ctx) { return ctx ? (ctx->kids.empty() ? 1 : (int)ctx->kids.size()) : 0; }

// the parent handle: shapes and the claim / evaluation bookkeeping of the one transcript; no device memory of its own
static int multi_tables_new(zksc_ctx* ctx, uint32_t n_vars, uint32_t B, uint32_t P, const uint32_t* degree, zksc_tables** out) {
    *out = nullptr;
    if (!degree) return ZKSC_ERR_SHAPE;
    if (B < 1 || P < 1 || P > ZKSC_MAX_PRODUCTS) FAIL(ZKSC_ERR_SHAPE, "need 1 <= n_products <= ZKSC_MAX_PRODUCTS and n_proofs >= 1");
    zksc_tables* t = new zksc_tables();
    t->ctx = ctx; t->n_vars = n_vars; t->B = B; t->P = P; t->Dtot = 0; t->E = 0;
    for (uint32_t p = 0; p < P; p++) {
        if (degree[p] < 1 || degree[p] > ZKSC_MAX_DEGREE) { delete t; FAIL(ZKSC_ERR_UNSUPPORTED, "product degree must be in 1..ZKSC_MAX_DEGREE"); }
        t->deg[p] = degree[p]; t->koff[p] = t->Dtot; t->eoff[p] = t->E;
        t->Dtot += degree[p]; t->E += degree[p] + 1;
    }
    t->n_local0 = 1ull << n_vars;
    t->vars_left = n_vars; t->cur_n = t->n_local0; t->where = 0; t->pending = false;
    *out = t;
    return ZKSC_OK;
}
static void multi_adopt(zksc_tables* t, std::vector<zksc_tables*>& kids) {
    for (zksc_tables* k : kids) k->raw = true;
    t->kids = kids;
}
static int multi_tables_synth(zksc_ctx* ctx, uint32_t n_vars, uint32_t B, uint32_t P, const uint32_t* degree, uint64_t seed, zksc_tables** out) {
    zksc_tables* t = nullptr;
    TRY(multi_tables_new(ctx, n_vars, B, P, degree, &t));
    std::vector<zksc_tables*> kids(ctx->kids.size(), nullptr);
    const int rc = for_each_kid(ctx, [&](uint32_t g) { return zksc_tables_synth(ctx->kids[g], n_vars, B, P, degree, seed, &kids[g]); });
    if (rc != ZKSC_OK) { for (zksc_tables* k : kids) zksc_tables_free(k); delete t; return rc; }
    multi_adopt(t, kids);
    *out = t;
    return ZKSC_OK;
}
// Every table crosses PCIe ONCE: it is staged chunk by chunk in device 0's memory, and every device picks the entries of its
// shard (index = rank mod G) out of that staging buffer over NVLink.
static int multi_tables_upload(zksc_ctx* ctx, uint32_t n_vars, uint32_t B, uint32_t P, const uint32_t* degree, const uint64_t* const* host_tables,
                               zksc_tables** out) {
    zksc_tables* t = nullptr;
    TRY(multi_tables_new(ctx, n_vars, B, P, degree, &t));
    const uint32_t G = (uint32_t)ctx->kids.size();
    std::vector<zksc_tables*> kids(G, nullptr);
    auto bail = [&](int rc, const std::string& why) {
        for (zksc_tables* k : kids) zksc_tables_free(k);
        delete t;
        ctx->err = why;
        return rc;
    };
    for (uint32_t g = 0; g < G; g++) {
        const int rc = tables_alloc(ctx->kids[g], n_vars, B, P, degree, &kids[g]);
        if (rc != ZKSC_OK) return bail(rc, ctx->kids[g]->err);
    }
    const uint64_t N = 1ull << n_vars, chunk = std::min<uint64_t>(N, 1ull << 22);
    zksc_ctx* k0 = ctx->kids[0];
    Fr* stage = nullptr;
    std::vector<cudaEvent_t> picked(G, nullptr);
    cudaEvent_t staged = nullptr;
    cudaError_t e = cudaSetDevice(k0->device);
    if (e == cudaSuccess) e = cudaMalloc(&stage, chunk * sizeof(Fr));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&staged, cudaEventDisableTiming);
    for (uint32_t g = 0; g < G && e == cudaSuccess; g++) {
        e = cudaSetDevice(ctx->kids[g]->device);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&picked[g], cudaEventDisableTiming);
    }
    const size_t n_tabs = (size_t)B * t->Dtot;
    bool first = true;
    for (size_t i = 0; i < n_tabs && e == cudaSuccess; i++)
        for (uint64_t c0 = 0; c0 < N && e == cudaSuccess; c0 += chunk) {
            e = cudaSetDevice(k0->device);
            for (uint32_t g = 0; g < G && e == cudaSuccess && !first; g++) e = cudaStreamWaitEvent(k0->stream, picked[g], 0);   // the buffer is free again
            if (e == cudaSuccess) e = cudaMemcpyAsync(stage, host_tables[i] + 4 * c0, chunk * sizeof(Fr), cudaMemcpyHostToDevice, k0->stream);
            if (e == cudaSuccess) e = cudaEventRecord(staged, k0->stream);
            for (uint32_t g = 0; g < G && e == cudaSuccess; g++) {
                zksc_ctx* k = ctx->kids[g];
                e = cudaSetDevice(k->device);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(k->stream, staged, 0);
                if (e != cudaSuccess) break;
                const uint64_t n_pick = chunk / G;
                pick_shard_kernel<<<grid_for(k, n_pick, 256, 8), 256, 0, k->stream>>>(stage, kids[g]->orig + i * kids[g]->n_local0 + c0 / G, n_pick, G, g);
                k->launches++;
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaEventRecord(picked[g], k->stream);
            }
            first = false;
        }
    for (uint32_t g = 0; g < G; g++) {
        cudaSetDevice(ctx->kids[g]->device);
        const cudaError_t e2 = cudaStreamSynchronize(ctx->kids[g]->stream);      // the host buffers are only borrowed for the call
        if (e == cudaSuccess) e = e2;
        if (picked[g]) cudaEventDestroy(picked[g]);
    }
    cudaSetDevice(k0->device);
    if (staged) cudaEventDestroy(staged);
    cudaFree(stage);
    if (e != cudaSuccess) return bail(e == cudaErrorMemoryAllocation ? ZKSC_ERR_OOM : ZKSC_ERR_CUDA, std::string("multi-GPU upload: ") + cudaGetErrorString(e));
    multi_adopt(t, kids);
    *out = t;
    return ZKSC_OK;
}
static int multi_tables_reset(zksc_tables* t) {
    for (zksc_tables* k : t->kids) { const int rc = zksc_tables_reset(k); if (rc != ZKSC_OK) { t->ctx->err = k->ctx->err; return rc; } }
    t->vars_left = t->n_vars;
    t->pending = false;
    t->last_evals_valid = false;
    t->claim_valid = false;
    return ZKSC_OK;
}
static int multi_tables_free(zksc_tables* t) {
    for (zksc_tables* k : t->kids) zksc_tables_free(k);
    delete t;
    return ZKSC_OK;
}
// Without a resident kernel (disabled, or degrees it does not take) the shards shrink to one entry per table and rank; the parent
// collects those G entries per table through the host and hands every device the whole residual table (replicated from then on).
static int multi_gather(zksc_tables* t) {
    zksc_ctx* ctx = t->ctx;
    const uint32_t G = (uint32_t)ctx->kids.size();
    const size_t nt = (size_t)t->B * t->Dtot;
    std::vector<Fr> mine(nt * G), all(nt * G);
    for (uint32_t g = 0; g < G; g++) {
        zksc_tables* k = t->kids[g];
        zksc_ctx* kc = k->ctx;
        CK(cudaSetDevice(kc->device));
        { const int rc = flush_pending(k); if (rc != ZKSC_OK) { ctx->err = kc->err; return rc; } }
        if (k->cur_n != 1) FAIL(ZKSC_ERR_STATE, "multi_gather: local tables not exhausted");
        Geo ge = geo_of(k, k->where);
        cudaError_t e = cudaMemcpy2DAsync(mine.data() + (size_t)g * nt, sizeof(Fr), ge.base, ge.tab_stride * sizeof(Fr), sizeof(Fr), nt, cudaMemcpyDeviceToHost, kc->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(kc->stream);
        if (e != cudaSuccess) FAIL(ZKSC_ERR_CUDA, std::string("multi_gather: ") + cudaGetErrorString(e));
    }
    for (size_t tab = 0; tab < nt; tab++)
        for (uint32_t g = 0; g < G; g++) all[tab * G + g] = mine[(size_t)g * nt + tab];     // entry index of the residual table == rank
    for (uint32_t g = 0; g < G; g++) {
        zksc_tables* k = t->kids[g];
        zksc_ctx* kc = k->ctx;
        CK(cudaSetDevice(kc->device));
        cudaError_t e = cudaMemcpy2DAsync(k->tail, k->tail_stride * sizeof(Fr), all.data(), (size_t)G * sizeof(Fr), (size_t)G * sizeof(Fr), nt, cudaMemcpyHostToDevice, kc->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(kc->stream);
        if (e != cudaSuccess) FAIL(ZKSC_ERR_CUDA, std::string("multi_gather: ") + cudaGetErrorString(e));
        k->where = 2;
        k->cur_n = G;
    }
    return ZKSC_OK;
}
static int multi_gather_if_due(zksc_tables* t) {
    const zksc_tables* k0 = t->kids[0];
    if (k0->tail_running || k0->where == 2) return ZKSC_OK;
    const uint64_t n_after = k0->pending ? k0->cur_n / 2 : k0->cur_n;
    return n_after == 1 ? multi_gather(t) : ZKSC_OK;
}
static int multi_round_evals(zksc_tables* t, uint64_t* out, uint32_t npts_cap) {
    zksc_ctx* ctx = t->ctx;
    if (t->vars_left == 0) FAIL(ZKSC_ERR_STATE, "all variables are bound");
    if (t->r0_valid && t->vars_left == t->n_vars && !t->pending && npts_cap > ZKSC_MAX_DEGREE) {
        memcpy(out, t->r0_cache.data(), t->r0_cache.size() * sizeof(uint64_t));
        t->r0_valid = false;
        t->last_evals = t->r0_cache;
        t->last_evals_valid = true;
        return ZKSC_OK;
    }
    GpuPoolSession session(ctx);
    TRY(multi_gather_if_due(t));
    const uint32_t G = (uint32_t)ctx->kids.size();
    const size_t n_res = (size_t)t->B * t->E;
    const bool fold = t->pending, full = npts_cap > ZKSC_MAX_DEGREE;
    const bool skip1 = fold && full && t->claim_valid;
    static thread_local std::vector<uint64_t> part_store;
    part_store.assign((size_t)G * n_res * 4, 0);
    uint64_t* const part = part_store.data();      // (the worker threads must not name the thread_local itself: they would get their own)
    TRY(for_each_kid(ctx, [&](uint32_t g) { return round_evals_impl(t->kids[g], part + (size_t)g * n_res * 4, npts_cap); }));
    // a sharded round: the G partial evaluations add up; a replicated round (after the gather): every device holds the whole value
    const bool partial = t->kids[0]->last_partial;
    for (size_t i = 0; i < n_res; i++) {
        FrH acc = load_h(&part[4 * i]);
        if (partial) for (uint32_t g = 1; g < G; g++) acc = host::add(acc, load_h(&part[((size_t)g * n_res + i) * 4]));
        store_h(out + 4 * i, acc);
    }
    if (fold) t->pending = false;
    finish_round(t, out, skip1, full, npts_cap);
    return ZKSC_OK;
}
static int multi_bind(zksc_tables* t, const uint64_t* challenges) {
    zksc_ctx* ctx = t->ctx;
    if (t->vars_left == 0) FAIL(ZKSC_ERR_STATE, "all variables are bound");
    GpuPoolSession session(ctx);
    const bool claims = t->last_evals_valid;
    if (claims) {
        t->claim.resize((size_t)t->B * t->P);
        for_each_proof(ctx, t->B, [&](uint32_t b) {
            std::vector<FrH> ys;
            for (uint32_t p = 0; p < t->P; p++) {
                ys.clear();
                for (uint32_t i = 0; i <= t->deg[p]; i++) ys.push_back(load_h(&t->last_evals[((size_t)b * t->E + t->eoff[p] + i) * 4]));
                t->claim[(size_t)b * t->P + p] = host::SparseUnivariatePolynomial::evaluate_evals_at(ys, load_h(challenges + 4 * b));
            }
        });
    }
    // a bind on top of an unapplied one flushes the first (stand-alone fold) -- which may be what exhausts the shards
    if (t->kids[0]->pending && !t->kids[0]->tail_running) {
        const zksc_tables* k0 = t->kids[0];
        if (k0->where != 2 && k0->cur_n / 2 == 1) TRY(multi_gather(t));
    }
    if (!t->kids[0]->pending && !t->kids[0]->tail_running && t->kids[0]->where != 2 && t->kids[0]->cur_n == 1) TRY(multi_gather(t));
    TRY(for_each_kid(ctx, [&](uint32_t g) {
        t->kids[g]->raw_claim_next = claims;
        return zksc_bind(t->kids[g], challenges);
    }));
    t->pending = true;
    t->vars_left--;
    t->claim_valid = claims;
    t->last_evals_valid = false;
    return ZKSC_OK;
}
static int multi_residual(zksc_tables* t, uint64_t* out) {
    zksc_ctx* ctx = t->ctx;
    for (zksc_tables* k : t->kids) { const int rc = tail_stop(k); if (rc != ZKSC_OK) { ctx->err = k->ctx->err; return rc; } }
    TRY(multi_gather_if_due(t));
    if (t->kids[0]->where != 2) FAIL(ZKSC_ERR_UNSUPPORTED, "sharded residual is only available once each device holds one entry per table");
    for (zksc_tables* k : t->kids) {
        CK(cudaSetDevice(k->ctx->device));
        const int rc = flush_pending(k);
        if (rc != ZKSC_OK) { ctx->err = k->ctx->err; return rc; }
    }
    t->pending = false;
    const int rc = zksc_residual(t->kids[0], out);
    if (rc != ZKSC_OK) ctx->err = t->kids[0]->ctx->err;
    return rc;
}

extern "C" int zksc_poly_sum(zksc_tables* t, uint64_t* out) {
    if (!t || !out) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (t->vars_left != t->n_vars || t->pending) FAIL(ZKSC_ERR_STATE, "zksc_poly_sum needs unbound tables");
    GpuPoolSession gpu_session(t->kids.empty() ? nullptr : ctx);
    std::vector<uint64_t> ev((size_t)t->B * t->E * 4);
    if (t->n_vars == 0) {
        // a single entry per table: the sum is the product itself
        std::vector<uint64_t> resid((size_t)t->B * t->Dtot * 4);
        TRY(zksc_residual(t, resid.data()));
        for (uint32_t b = 0; b < t->B; b++) {
            FrH s = host::kZero;
            for (uint32_t p = 0; p < t->P; p++) {
                FrH prod = host::kOne;
                for (uint32_t k = 0; k < t->deg[p]; k++) prod = host::mul(prod, load_h(&resid[((size_t)b * t->Dtot + t->koff[p] + k) * 4]));
                s = host::add(s, prod);
            }
            store_h(out + 4 * b, s);
        }
        return ZKSC_OK;
    }
    t->r0_valid = false;
    TRY(round_evals_impl(t, ev.data(), ZKSC_MAX_DEGREE + 1));
    for (uint32_t b = 0; b < t->B; b++) {
        FrH s = host::kZero;
        for (uint32_t p = 0; p < t->P; p++) {
            const uint64_t* e = &ev[((size_t)b * t->E + t->eoff[p]) * 4];
            s = host::add(s, host::add(load_h(e), load_h(e + 4)));
        }
        store_h(out + 4 * b, s);
    }
    zksc_tables_reset(t);
    t->r0_cache = ev;
    t->r0_valid = true;
    return ZKSC_OK;
}

// multi-GPU handle: every device converts its shard, the host interleaves the 32-byte elements (entry i sits on device i mod G at i / G)
static int multi_tables_to_bytes(zksc_tables* t, uint32_t proof, uint8_t* out) {
    zksc_ctx* ctx = t->ctx;
    if (proof >= t->B) FAIL(ZKSC_ERR_SHAPE, "proof index");
    const uint32_t G = (uint32_t)t->kids.size();
    const uint64_t NL = t->kids[0]->n_local0;
    std::vector<uint8_t> shard((size_t)NL * 32);
    for (uint32_t g = 0; g < G; g++) {
        zksc_tables* k = t->kids[g];
        zksc_ctx* kc = k->ctx;
        CK(cudaSetDevice(kc->device));
        { const int rc = quiesce(kc); if (rc != ZKSC_OK) { ctx->err = kc->err; return rc; } }
        Fr* tmp = nullptr;
        CK(cudaMallocAsync((void**)&tmp, NL * sizeof(Fr), kc->stream));
        for (uint32_t tab = 0; tab < t->Dtot; tab++) {
            to_bytes_kernel<<<grid_for(kc, NL, 256, 8), 256, 0, kc->stream>>>(k->orig + ((size_t)proof * t->Dtot + tab) * NL, tmp, NL);
            kc->launches++;
            cudaError_t e = cudaMemcpyAsync(shard.data(), tmp, NL * sizeof(Fr), cudaMemcpyDeviceToHost, kc->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(kc->stream);
            if (e != cudaSuccess) { cudaFreeAsync(tmp, kc->stream); FAIL(ZKSC_ERR_CUDA, std::string("to_bytes: ") + cudaGetErrorString(e)); }
            uint8_t* dst = out + (size_t)tab * NL * G * 32;
            for (uint64_t i = 0; i < NL; i++) memcpy(dst + (i * G + g) * 32, shard.data() + i * 32, 32);
        }
        cudaFreeAsync(tmp, kc->stream);
    }
    return ZKSC_OK;
}

extern "C" int zksc_tables_to_bytes(zksc_tables* t, uint32_t proof, uint8_t* out) {
    if (!t || !out) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (!t->kids.empty()) return multi_tables_to_bytes(t, proof, out);
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    if (ctx->n_ranks != 1) FAIL(ZKSC_ERR_UNSUPPORTED, "zksc_tables_to_bytes: single-rank contexts only");
    if (proof >= t->B) FAIL(ZKSC_ERR_SHAPE, "proof index");
    const uint64_t N = t->n_local0;
    const uint64_t chunk = N < (1ull << 22) ? N : (1ull << 22);
    DevBuf tmpbuf(ctx);
    CK(dev_alloc(ctx, (void**)&tmpbuf.p, chunk * sizeof(Fr)));
    Fr* const tmp = tmpbuf.p;
    for (uint32_t k = 0; k < t->Dtot; k++) {
        const Fr* src = t->orig + ((size_t)proof * t->Dtot + k) * N;
        for (uint64_t off = 0; off < N; off += chunk) {
            to_bytes_kernel<<<grid_for(ctx, chunk, 256, 8), 256, 0, ctx->stream>>>(src + off, tmp, chunk);
            ctx->launches++;
            cudaError_t e = cudaMemcpyAsync(out + ((size_t)k * N + off) * 32, tmp, chunk * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { ctx->err = std::string("to_bytes: ") + cudaGetErrorString(e); return ZKSC_ERR_CUDA; }
        }
    }
    return ZKSC_OK;
}

// ------------------------------------------------------------------------------------------------
// provers: host round loop + transcript (sumcheck.rs:29-61, composed_sumcheck.rs:32-67,
// multi_composed_sumcheck.rs:47-120)
// ------------------------------------------------------------------------------------------------
extern "C" uint32_t zksc_msg_stride(int protocol, uint32_t n_products, const uint32_t* degree) {
    uint32_t dmax = 0;
    for (uint32_t p = 0; p < n_products; p++) dmax = degree[p] > dmax ? degree[p] : dmax;
    if (protocol == ZKSC_PROTO_SUMCHECK) return 2;
    if (protocol == ZKSC_PROTO_COMPOSED) return dmax + 1;
    return 2 * (dmax + 1);
}

// The round loop of the three provers (multi_composed_sumcheck.rs:75-107, sumcheck.rs:38-57, composed_sumcheck.rs:44-63) over ALL variables of
// the handle, on transcripts the caller has prepared.  The messages and challenges of round j of the handle go to slot round0 + j of a proof
// of n rounds in total: a sumcheck whose tables change form half-way (gkr_linear.cuh) is two such runs on one transcript.
static int prove_run(zksc_tables* t, int protocol, std::vector<host::FiatShamirTranscript>& tr, uint32_t n, uint32_t round0, uint32_t stride,
                     uint64_t* round_msgs, uint32_t* round_len, uint64_t* challenges) {
    zksc_ctx* ctx = t->ctx;
    const uint32_t B = t->B;
    std::vector<uint64_t> ev((size_t)B * t->E * 4), chal((size_t)B * 4);
    // many independent proofs per round: their transcripts go to a few host threads (they spin only during this call)
    if (B >= kPoolMinProofs && !ctx->pool) {
        const char* e = getenv("ZKSC_HOST_THREADS");
        int want = e ? atoi(e) : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
        if (want > 1) ctx->pool = new HostPool(want - 1);
    }
    struct PoolSession {
        zksc_ctx* c;
        explicit PoolSession(zksc_ctx* ctx, bool on) : c(on ? ctx : nullptr) { if (c) { c->pool->begin(); c->pool_active = true; } }
        ~PoolSession() { if (c) { c->pool_active = false; c->pool->end(); } }
    } pool_session(ctx, ctx->pool != nullptr && B >= kPoolMinProofs);
    struct InProve { zksc_tables* t; explicit InProve(zksc_tables* x) : t(x) { t->in_prove = true; } ~InProve() { t->in_prove = false; } } in_prove(t);
    for (uint32_t round = round0; round < round0 + t->n_vars; round++) {
        const auto p0 = std::chrono::steady_clock::now();
        TRY(round_evals_impl(t, ev.data(), ZKSC_MAX_DEGREE + 1));
        const auto p1 = std::chrono::steady_clock::now();
        for_each_proof(ctx, B, [&](uint32_t b) {
            uint64_t* msg = round_msgs + ((size_t)b * n + round) * stride * 4;
            uint32_t* len = round_len + (size_t)b * n + round;
            const uint64_t* e = &ev[(size_t)b * t->E * 4];
            static thread_local std::vector<uint8_t> bytes;
            bytes.clear();
            if (protocol == ZKSC_PROTO_SUMCHECK || protocol == ZKSC_PROTO_COMPOSED) {
                const uint32_t cnt = t->deg[0] + 1;
                memcpy(msg, e, (size_t)cnt * 32);
                *len = cnt;
                for (uint32_t i = 0; i < cnt; i++) {
                    uint8_t be[32];
                    host::to_be_bytes(load_h(e + 4 * i), be);
                    bytes.insert(bytes.end(), be, be + 32);
                }
            } else {
                // round_poly = sum over the products of interpolate(evaluations)  (:77, :91-94).  SparseUnivariatePolynomial's interpolation
                // drops the monomials whose coefficient is zero, its Add merges by power and KEEPS a sum that happens to be zero
                // (sparse_univariate.rs:52-60, :159-203): a power is present iff some product has a nonzero coefficient there, and carries
                // the sum of all products' coefficients -- formed here on the dense coefficients, without building the sparse objects.
                FrH coeff[ZKSC_MAX_DEGREE + 1];
                bool present[ZKSC_MAX_DEGREE + 1] = {};
                static thread_local std::vector<FrH> ys;
                for (uint32_t p = 0; p < t->P; p++) {
                    ys.clear();
                    for (uint32_t i = 0; i <= t->deg[p]; i++) ys.push_back(load_h(e + 4 * (t->eoff[p] + i)));
                    const std::vector<FrH> dense = host::SparseUnivariatePolynomial::dense_interpolate_evals(ys);
                    for (size_t k = 0; k < dense.size(); k++) {
                        if (dense[k] == host::kZero) continue;
                        coeff[k] = present[k] ? host::add(coeff[k], dense[k]) : dense[k];
                        present[k] = true;
                    }
                }
                uint32_t nm = 0;
                for (uint32_t k = 0; k <= ZKSC_MAX_DEGREE; k++) {
                    if (!present[k]) continue;
                    const FrH& pw = host::SparseUnivariatePolynomial::small_mont(k);
                    store_h(msg + 8 * nm, coeff[k]);
                    store_h(msg + 8 * nm + 4, pw);
                    uint8_t b64[64];                                   // SparseUnivariatePolynomial::to_bytes (:27-34)
                    host::to_be_bytes(coeff[k], b64);
                    host::to_be_bytes(pw, b64 + 32);
                    bytes.insert(bytes.end(), b64, b64 + 64);
                    nm++;
                }
                *len = nm;
            }
            tr[b].commit(bytes);                                      // :97
            FrH r = tr[b].evaluate_challenge_into_field();            // :99
            store_h(&chal[4 * b], r);
            store_h(challenges + ((size_t)b * n + round) * 4, r);
        });
        const auto p2 = std::chrono::steady_clock::now();
        TRY(zksc_bind(t, chal.data()));                               // :103-105 (deferred, fused)
        const auto p3 = std::chrono::steady_clock::now();
        if (round < ctx->round_us.size()) ctx->round_us[round] = std::chrono::duration<double, std::micro>(p3 - p0).count();
        if (ctx->profile) {
            auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
            fprintf(stderr, "[zksc profile] round %2u: evals %7.1f us (launch %6.1f, wait %7.1f)  transcript %5.1f  bind %5.1f\n", round, us(p0, p1),
                    ctx->prof_launch, ctx->prof_wait, us(p1, p2), us(p2, p3));
        }
    }
    return ZKSC_OK;
}

extern "C" int zksc_prove(zksc_tables* t, int protocol, const uint64_t* sums, uint64_t* round_msgs, uint32_t* round_len, uint64_t* challenges) {
    if (!t || !round_msgs || !round_len || !challenges) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = t->ctx;
    if (protocol < ZKSC_PROTO_SUMCHECK || protocol > ZKSC_PROTO_MULTI_FULL) FAIL(ZKSC_ERR_SHAPE, "unknown protocol");
    if (protocol == ZKSC_PROTO_SUMCHECK && (t->P != 1 || t->deg[0] != 1)) FAIL(ZKSC_ERR_SHAPE, "Sumcheck::prove takes one multilinear table");
    if (protocol == ZKSC_PROTO_COMPOSED && t->P != 1) FAIL(ZKSC_ERR_SHAPE, "ComposedSumcheck::prove takes one product");
    if (protocol != ZKSC_PROTO_COMPOSED && !sums) FAIL(ZKSC_ERR_SHAPE, "sums is NULL");
    if (t->vars_left != t->n_vars || t->pending) FAIL(ZKSC_ERR_STATE, "tables are partially bound; call zksc_tables_reset first");
    const uint32_t stride = zksc_msg_stride(protocol, t->P, t->deg);
    const uint32_t B = t->B, n = t->n_vars;
    GpuPoolSession gpu_session(t->kids.empty() ? nullptr : ctx);
    std::vector<host::FiatShamirTranscript> tr(B);
    if (protocol == ZKSC_PROTO_MULTI_FULL) {
        // transcript.commit(&composed_poly_to_bytes(&poly))  multi_composed_sumcheck.rs:52
        if (ctx->n_ranks != 1) FAIL(ZKSC_ERR_UNSUPPORTED, "MULTI_FULL (absorbs every table entry) is single-rank only; use MULTI_PARTIAL");
        std::vector<uint8_t> bytes((size_t)t->Dtot * t->n_local0 * 32);
        for (uint32_t b = 0; b < B; b++) {
            TRY(zksc_tables_to_bytes(t, b, bytes.data()));
            tr[b].commit(bytes);
        }
    }
    if (protocol != ZKSC_PROTO_COMPOSED)
        for (uint32_t b = 0; b < B; b++) tr[b].commit_field(load_h(sums + 4 * b));  // :70 / sumcheck.rs:34-35

    memset(round_msgs, 0, (size_t)B * n * stride * 32);
    ctx->round_us.assign(n, 0.0);
    return prove_run(t, protocol, tr, n, 0, stride, round_msgs, round_len, challenges);
}

extern "C" int zksc_proof_to_bytes(int protocol, uint32_t n_vars, uint32_t msg_stride, const uint64_t* round_msgs, const uint32_t* round_len,
                                   uint8_t* out, size_t* out_len) {
    if (!out_len || (n_vars && (!round_msgs || !round_len))) return ZKSC_ERR_SHAPE;
    const uint32_t per = (protocol == ZKSC_PROTO_MULTI_PARTIAL || protocol == ZKSC_PROTO_MULTI_FULL) ? 2 : 1;
    size_t n = 0;
    for (uint32_t r = 0; r < n_vars; r++) {
        if ((uint64_t)round_len[r] * per > msg_stride) return ZKSC_ERR_SHAPE;   // a round message never exceeds its slot
        for (uint32_t i = 0; i < round_len[r] * per; i++) {
            if (out) host::to_be_bytes(load_h(round_msgs + ((size_t)r * msg_stride + i) * 4), out + n);
            n += 32;
        }
    }
    *out_len = n;
    return ZKSC_OK;
}

extern "C" int zksc_verify_rounds(int protocol, uint32_t n_vars, uint32_t msg_stride, const uint64_t* sum, const uint64_t* round_msgs,
                                  const uint32_t* round_len, const uint8_t* absorbed_prefix, size_t prefix_len, uint64_t* subclaim_sum,
                                  uint64_t* challenges) {
    if (!sum || !subclaim_sum || !challenges || (n_vars && (!round_msgs || !round_len))) return ZKSC_ERR_SHAPE;
    if (protocol < ZKSC_PROTO_SUMCHECK || protocol > ZKSC_PROTO_MULTI_FULL) return ZKSC_ERR_SHAPE;
    {   // the proof is untrusted input: every round message must fit the slot the caller's stride gives it
        const uint32_t per = (protocol == ZKSC_PROTO_MULTI_PARTIAL || protocol == ZKSC_PROTO_MULTI_FULL) ? 2 : 1;
        for (uint32_t r = 0; r < n_vars; r++)
            if ((uint64_t)round_len[r] * per > msg_stride) return ZKSC_ERR_SHAPE;
    }
    host::FiatShamirTranscript tr;
    if (absorbed_prefix && prefix_len) tr.commit(absorbed_prefix, prefix_len);
    FrH claimed = load_h(sum);
    if (protocol != ZKSC_PROTO_COMPOSED) tr.commit_field(claimed);
    std::vector<uint8_t> bytes;
    for (uint32_t r = 0; r < n_vars; r++) {
        const uint64_t* msg = round_msgs + (size_t)r * msg_stride * 4;
        bytes.clear();
        FrH p0, p1, c;
        if (protocol == ZKSC_PROTO_SUMCHECK) {
            // sumcheck.rs:73-92: check, then commit, then challenge
            if (round_len[r] != 2) return ZKSC_ERR_SHAPE;
            FrH h0 = load_h(msg), h1 = load_h(msg + 4);
            if (host::add(h0, h1) != claimed) return ZKSC_ERR_VERIFY;
            uint8_t be[64];
            host::to_be_bytes(h0, be); host::to_be_bytes(h1, be + 32);
            tr.commit(be, 64);
            c = tr.evaluate_challenge_into_field();
            claimed = host::add(host::mul(c, h1), host::mul(host::sub(host::kOne, c), h0));  // uni_poly.evaluation(&[c])
        } else {
            host::SparseUnivariatePolynomial poly;
            if (protocol == ZKSC_PROTO_COMPOSED) {
                std::vector<FrH> ys;
                for (uint32_t i = 0; i < round_len[r]; i++) {
                    ys.push_back(load_h(msg + 4 * i));
                    uint8_t be[32];
                    host::to_be_bytes(ys.back(), be);
                    bytes.insert(bytes.end(), be, be + 32);
                }
                poly = host::SparseUnivariatePolynomial::interpolate_evals(ys);  // composed_sumcheck.rs:81-83
            } else {
                for (uint32_t i = 0; i < round_len[r]; i++) poly.monomial.push_back({load_h(msg + 8 * i), load_h(msg + 8 * i + 4)});
                poly.to_bytes(bytes);
            }
            tr.commit(bytes);
            c = tr.evaluate_challenge_into_field();
            p0 = poly.evaluate(host::kZero);
            p1 = poly.evaluate(host::kOne);
            if (host::add(p0, p1) != claimed) return ZKSC_ERR_VERIFY;
            claimed = poly.evaluate(c);
        }
        store_h(challenges + 4 * r, c);
    }
    store_h(subclaim_sum, claimed);
    return ZKSC_OK;
}

extern "C" int zksc_evaluate(zksc_tables* t, const uint64_t* points, uint64_t* out) {
    if (!t || !points || !out) return ZKSC_ERR_STATE;
    GpuPoolSession gpu_session(t->kids.empty() ? nullptr : t->ctx);
    zksc_tables_reset(t);
    std::vector<uint64_t> chal((size_t)t->B * 4);
    for (uint32_t j = 0; j < t->n_vars; j++) {
        for (uint32_t b = 0; b < t->B; b++) memcpy(&chal[4 * b], points + ((size_t)b * t->n_vars + j) * 4, 32);
        TRY(zksc_bind(t, chal.data()));
    }
    std::vector<uint64_t> resid((size_t)t->B * t->Dtot * 4);
    TRY(zksc_residual(t, resid.data()));
    for (uint32_t b = 0; b < t->B; b++) {
        FrH s = host::kZero;
        for (uint32_t p = 0; p < t->P; p++) {
            FrH prod = host::kOne;  // ComposedMultilinear::evaluation  composed_multilinear.rs:52-61
            for (uint32_t k = 0; k < t->deg[p]; k++) prod = host::mul(prod, load_h(&resid[((size_t)b * t->Dtot + t->koff[p] + k) * 4]));
            s = host::add(s, prod);
        }
        store_h(out + 4 * b, s);
    }
    zksc_tables_reset(t);
    return ZKSC_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone Multilinear operations
// ------------------------------------------------------------------------------------------------

extern "C" int zksc_ml_partial_evaluation(zksc_ctx* ctx, const uint64_t* evals, uint64_t n, const uint64_t* r, uint32_t variable_index, uint64_t* out) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) ctx = ctx->kids[0];      // stand-alone vector operations run on the first device
    if (!evals || !r || !out) FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    if (n < 2 || (n & (n - 1))) FAIL(ZKSC_ERR_SHAPE, "Number of evaluations must be a power of 2 (and at least 2 to bind a variable)");
    uint32_t n_vars_ml = 0;
    while ((1ull << n_vars_ml) < n) n_vars_ml++;
    // the reference asserts variable_index < n/2 (polynomial/src/utils.rs:30-34); an index past the last variable would pair nothing
    // (and a shift by >= 64 would be undefined), so it is rejected before any shift
    if (variable_index >= n_vars_ml) FAIL(ZKSC_ERR_SHAPE, "variable_index must be less than n/2 and name an existing variable");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    DevBuf in(ctx), o(ctx);
    CK(dev_alloc(ctx, (void**)&in.p, n * sizeof(Fr)));
    CK(dev_alloc(ctx, (void**)&o.p, n / 2 * sizeof(Fr)));
    CK(cudaMemcpyAsync(in.p, evals, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    FoldArgs a;
    a.in = in.p; a.out = o.p;
    a.in_tab_stride = a.in_proof_stride = a.out_tab_stride = a.out_proof_stride = 0;
    a.n_out = n / 2; a.s = n >> (variable_index + 1); a.n_tabs = 1;
    memcpy(a.chal[0].l, r, 32);
    fold_kernel<<<dim3(grid_for(ctx, a.n_out, 256, 8), 1), 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, o.p, n / 2 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

extern "C" int zksc_ml_evaluation(zksc_ctx* ctx, const uint64_t* evals, uint64_t n, const uint64_t* points, uint32_t n_points, uint64_t* out) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) ctx = ctx->kids[0];      // stand-alone vector operations run on the first device
    if (!evals || !out || (n_points && !points)) FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    if (n < 1 || (n & (n - 1))) FAIL(ZKSC_ERR_SHAPE, "Number of evaluations must be a power of 2");
    if ((1ull << n_points) != n) FAIL(ZKSC_ERR_SHAPE, "Number of evaluation points must match the number of variables");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    DevBuf buf(ctx);
    CK(dev_alloc(ctx, (void**)&buf.p, n * sizeof(Fr)));
    CK(cudaMemcpyAsync(buf.p, evals, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    uint64_t cur = n;
    for (uint32_t j = 0; j < n_points; j++) {
        FoldArgs a;
        a.in = buf.p; a.out = buf.p;  // variable 0, in place
        a.in_tab_stride = a.in_proof_stride = a.out_tab_stride = a.out_proof_stride = 0;
        a.n_out = cur / 2; a.s = cur / 2; a.n_tabs = 1;
        memcpy(a.chal[0].l, points + 4 * j, 32);
        fold_kernel<<<dim3(grid_for(ctx, a.n_out, 256, 8), 1), 256, 0, ctx->stream>>>(a);
        ctx->launches++;
        cur /= 2;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, buf.p, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

extern "C" int zksc_ml_outer(zksc_ctx* ctx, int mul, const uint64_t* a, uint64_t na, const uint64_t* b, uint64_t nb, uint64_t* out) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) ctx = ctx->kids[0];      // stand-alone vector operations run on the first device
    if (!a || !b || !out || !na || !nb) FAIL(ZKSC_ERR_SHAPE, "empty operand");
    unsigned long long n = 0;
    if (__builtin_umulll_overflow(na, nb, &n) || (n & (n - 1))) FAIL(ZKSC_ERR_SHAPE, "Number of evaluations must be a power of 2");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    DevBuf da(ctx), db(ctx), dout(ctx);
    CK(dev_alloc(ctx, (void**)&da.p, na * sizeof(Fr)));
    CK(dev_alloc(ctx, (void**)&db.p, nb * sizeof(Fr)));
    CK(dev_alloc(ctx, (void**)&dout.p, n * sizeof(Fr)));
    CK(cudaMemcpyAsync(da.p, a, na * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(db.p, b, nb * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    outer_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(mul, da.p, na, db.p, nb, dout.p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

extern "C" int zksc_ml_elementwise(zksc_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) ctx = ctx->kids[0];      // stand-alone vector operations run on the first device
    if (!a || !b || !out || !n) FAIL(ZKSC_ERR_SHAPE, "empty operand");
    if (op < 0 || op > 3) FAIL(ZKSC_ERR_SHAPE, "op");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    DevBuf da(ctx), db(ctx), dout(ctx);
    const uint64_t nb = (op == EW_SCALE) ? 1 : n;
    CK(dev_alloc(ctx, (void**)&da.p, n * sizeof(Fr)));
    CK(dev_alloc(ctx, (void**)&db.p, nb * sizeof(Fr)));
    CK(dev_alloc(ctx, (void**)&dout.p, n * sizeof(Fr)));
    CK(cudaMemcpyAsync(da.p, a, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(db.p, b, nb * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    ew_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(op, da.p, db.p, dout.p, n);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}


// ------------------------------------------------------------------------------------------------
// multilinear KZG over BLS12-381 G1: commitment and opening proofs
// ------------------------------------------------------------------------------------------------
// sum_i scalars[i mod period] * points[i] on the device (g1.cuh); scalars and points already in device memory
static int g1_msm_device(zksc_ctx* ctx, const Fr* d_scalars, unsigned long long period, const g1::Jac* d_points, unsigned long long n, g1::Jac* d_out) {
    uint8_t* digits = nullptr;
    g1::Jac *buckets = nullptr, *windows = nullptr;
    CK(cudaMallocAsync((void**)&digits, (size_t)g1::kWindows * n, ctx->stream));
    CK(cudaMallocAsync((void**)&buckets, (size_t)g1::kWindows * g1::kBuckets * sizeof(g1::Jac), ctx->stream));
    CK(cudaMallocAsync((void**)&windows, (size_t)g1::kWindows * sizeof(g1::Jac), ctx->stream));
    g1::msm_digits_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(d_scalars, n, period, digits);
    g1::msm_bucket_kernel<<<g1::kWindows * 256 / 128, 128, 0, ctx->stream>>>(digits, d_points, n, buckets);
    g1::msm_window_kernel<<<1, g1::kWindows, 0, ctx->stream>>>(buckets, windows);
    g1::msm_combine_kernel<<<1, 1, 0, ctx->stream>>>(windows, d_out);
    ctx->launches += 4;
    CK(cudaGetLastError());
    cudaFreeAsync(digits, ctx->stream); cudaFreeAsync(buckets, ctx->stream); cudaFreeAsync(windows, ctx->stream);
    return ZKSC_OK;
}

extern "C" int zksc_g1_msm(zksc_ctx* ctx, const uint64_t* scalars, const uint64_t* points, uint64_t n, uint64_t* out) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) ctx = ctx->kids[0];
    if (!scalars || !points || !out || n == 0) FAIL(ZKSC_ERR_SHAPE, "empty operand");
    static_assert(sizeof(g1::Jac) == 144, "ark-ec Projective<g1::Config> is 18 x u64");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    Fr* ds = nullptr;
    g1::Jac *dp = nullptr, *dout = nullptr;
    CK(cudaMallocAsync((void**)&ds, n * sizeof(Fr), ctx->stream));
    CK(cudaMallocAsync((void**)&dp, n * sizeof(g1::Jac), ctx->stream));
    CK(cudaMallocAsync((void**)&dout, sizeof(g1::Jac), ctx->stream));
    CK(cudaMemcpyAsync(ds, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dp, points, n * sizeof(g1::Jac), cudaMemcpyHostToDevice, ctx->stream));
    int rc = g1_msm_device(ctx, ds, 0, dp, n, dout);
    if (rc == ZKSC_OK) {
        cudaError_t e = cudaMemcpyAsync(out, dout, sizeof(g1::Jac), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("zksc_g1_msm: ") + cudaGetErrorString(e); rc = ZKSC_ERR_CUDA; }
    }
    cudaFreeAsync(ds, ctx->stream); cudaFreeAsync(dp, ctx->stream); cudaFreeAsync(dout, ctx->stream);
    return rc;
}

// q = f(1, .) - f(0, .): second half minus first half (get_poly_quotient, kzg/src/utils.rs:12-17)
__global__ void __launch_bounds__(256) kzg_quotient_kernel(const Fr* f, Fr* q, unsigned long long half) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) st256(q + i, fr_sub(ld256(f + half + i), ld256(f + i)));
}

extern "C" int zksc_kzg_open(zksc_ctx* ctx, const uint64_t* evals, uint32_t n_vars, const uint64_t* points, const uint64_t* srs_g1, uint64_t* out_evaluation,
                             uint64_t* out_proofs) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) ctx = ctx->kids[0];
    if (!evals || !points || !srs_g1 || !out_evaluation || !out_proofs) FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    if (n_vars < 1 || n_vars > 30) FAIL(ZKSC_ERR_SHAPE, "zksc_kzg_open: 1 <= n_vars <= 30");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    const unsigned long long N = 1ull << n_vars;
    Fr *poly = nullptr, *quot = nullptr;
    g1::Jac *srs = nullptr, *proofs = nullptr;
    CK(cudaMallocAsync((void**)&poly, N * sizeof(Fr), ctx->stream));
    CK(cudaMallocAsync((void**)&quot, std::max<unsigned long long>(N / 2, 2) * sizeof(Fr), ctx->stream));
    CK(cudaMallocAsync((void**)&srs, N * sizeof(g1::Jac), ctx->stream));
    CK(cudaMallocAsync((void**)&proofs, (size_t)n_vars * sizeof(g1::Jac), ctx->stream));
    CK(cudaMemcpyAsync(poly, evals, N * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(srs, srs_g1, N * sizeof(g1::Jac), cudaMemcpyHostToDevice, ctx->stream));
    int rc = ZKSC_OK;
    unsigned long long cur = N;
    // multilinear_kzg.rs:58-82: per variable, quotient = f(1,.) - f(0,.), blown up to all n variables by repetition
    // (add_to_front / duplicate_evaluation: a periodic scalar vector, so the MSM just indexes the quotient modulo its length),
    // committed; then the polynomial is replaced by its remainder = partial_evaluation(point, 0)
    for (uint32_t v = 0; v < n_vars && rc == ZKSC_OK; v++) {
        const unsigned long long half = cur / 2;
        kzg_quotient_kernel<<<grid_for(ctx, half, 256, 8), 256, 0, ctx->stream>>>(poly, quot, half);
        ctx->launches++;
        rc = g1_msm_device(ctx, quot, half, srs, N, proofs + v);
        if (rc != ZKSC_OK) break;
        FoldArgs a;
        a.in = poly; a.out = poly;      // variable 0, in place (the remainder; for the last variable: the evaluation itself)
        a.in_tab_stride = a.in_proof_stride = a.out_tab_stride = a.out_proof_stride = 0;
        a.n_out = half; a.s = half; a.n_tabs = 1;
        memcpy(a.chal[0].l, points + 4 * v, 32);
        fold_kernel<<<dim3(grid_for(ctx, a.n_out, 256, 8), 1), 256, 0, ctx->stream>>>(a);
        ctx->launches++;
        cur = half;
    }
    if (rc == ZKSC_OK) {
        cudaError_t e = cudaGetLastError();
        // evaluation = poly_.evaluation(points) (:56); the chain of remainders ends in the same value (the reference panics otherwise, :84-86)
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_evaluation, poly, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_proofs, proofs, (size_t)n_vars * sizeof(g1::Jac), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("zksc_kzg_open: ") + cudaGetErrorString(e); rc = ZKSC_ERR_CUDA; }
    }
    cudaFreeAsync(poly, ctx->stream); cudaFreeAsync(quot, ctx->stream); cudaFreeAsync(srs, ctx->stream); cudaFreeAsync(proofs, ctx->stream);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// GKR layer driver
// ------------------------------------------------------------------------------------------------
#include "gkr_driver.cuh"
#include "gkr_linear.cuh"

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
extern "C" void zksc_fr_from_u64(uint64_t x, uint64_t out[4]) { store_h(out, host::from_u64(x)); }
extern "C" void zksc_fr_from_canonical(const uint64_t c[4], uint64_t out[4]) { store_h(out, host::from_canonical(c)); }
extern "C" void zksc_fr_to_canonical(const uint64_t m[4], uint64_t out[4]) { host::to_canonical(load_h(m), out); }
extern "C" void zksc_fr_from_canonical_batch(const uint64_t* c, uint64_t n, uint64_t* out) {
    for (uint64_t i = 0; i < n; i++) store_h(out + 4 * i, host::from_canonical(c + 4 * i));
}
extern "C" void zksc_fr_to_canonical_batch(const uint64_t* m, uint64_t n, uint64_t* out) {
    for (uint64_t i = 0; i < n; i++) host::to_canonical(load_h(m + 4 * i), out + 4 * i);
}
extern "C" void zksc_fr_to_be_bytes(const uint64_t m[4], uint8_t out[32]) { host::to_be_bytes(load_h(m), out); }
extern "C" void zksc_fr_from_be_bytes_mod_order(const uint8_t in[32], uint64_t out[4]) { store_h(out, host::from_be_bytes_mod_order(in)); }
extern "C" void zksc_fr_add(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { store_h(out, host::add(load_h(a), load_h(b))); }
extern "C" void zksc_fr_sub(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { store_h(out, host::sub(load_h(a), load_h(b))); }
extern "C" void zksc_fr_mul(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) { store_h(out, host::mul(load_h(a), load_h(b))); }

struct zksc_transcript {
    host::FiatShamirTranscript t;
};
extern "C" zksc_transcript* zksc_transcript_new(void) { return new zksc_transcript(); }
extern "C" void zksc_transcript_free(zksc_transcript* t) { delete t; }
extern "C" void zksc_transcript_commit(zksc_transcript* t, const uint8_t* data, size_t len) { t->t.commit(data, len); }
extern "C" void zksc_transcript_challenge(zksc_transcript* t, uint8_t out[32]) { t->t.challenge(out); }
extern "C" void zksc_transcript_challenge_field(zksc_transcript* t, uint64_t out[4]) { store_h(out, t->t.evaluate_challenge_into_field()); }

static uint32_t store_poly(const host::SparseUnivariatePolynomial& p, uint64_t* out) {
    for (size_t m = 0; m < p.monomial.size(); m++) {
        store_h(out + 8 * m, p.monomial[m].coeff);
        store_h(out + 8 * m + 4, p.monomial[m].pow);
    }
    return (uint32_t)p.monomial.size();
}
static host::SparseUnivariatePolynomial load_poly(const uint64_t* mono, uint32_t n) {
    host::SparseUnivariatePolynomial p;
    for (uint32_t i = 0; i < n; i++) p.monomial.push_back({load_h(mono + 8 * i), load_h(mono + 8 * i + 4)});
    return p;
}
extern "C" void zksc_round_slots_to_evals(uint32_t degree, uint64_t* values) { host::round_slots_to_evals(degree, values); }
extern "C" uint32_t zksc_sparse_interpolate(const uint64_t* ys, uint32_t n, uint64_t* out_mono) {
    std::vector<FrH> y;
    for (uint32_t i = 0; i < n; i++) y.push_back(load_h(ys + 4 * i));
    return store_poly(host::SparseUnivariatePolynomial::interpolate_evals(y), out_mono);
}
extern "C" uint32_t zksc_sparse_add(const uint64_t* a, uint32_t na, const uint64_t* b, uint32_t nb, uint64_t* out_mono) {
    return store_poly(load_poly(a, na) + load_poly(b, nb), out_mono);
}
extern "C" void zksc_sparse_evaluate(const uint64_t* mono, uint32_t n, const uint64_t point[4], uint64_t out[4]) {
    store_h(out, load_poly(mono, n).evaluate(load_h(point)));
}

static uint64_t h_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
extern "C" void zksc_synth_entry(uint64_t seed, uint64_t table, uint64_t index, uint64_t out[4]) {
    const uint64_t base = h_splitmix64(seed ^ h_splitmix64(table * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull));
    uint64_t c[4];
    for (int limb = 0; limb < 4; limb++) c[limb] = h_splitmix64(base + (index * 4 + limb) * 0x9E3779B97F4A7C15ull);
    store_h(out, host::from_canonical(c));
}


// ------------------------------------------------------------------------------------------------
// BLS12-381 pairing check on the host (host_pairing.hpp): the verifier half of the multilinear KZG / succinct GKR
// ------------------------------------------------------------------------------------------------
namespace {
namespace pr = zksc::pairing;
bool load_fq(const uint64_t* c, pr::Fq* out) {
    if (pr::geq_p(c)) return false;
    *out = pr::from_canonical(c);
    return true;
}
// affine points as canonical little-endian limbs; all-zero coordinates = the identity.  false: a coordinate >= p or a point off its curve.
bool load_g1(const uint64_t* c, pr::G1Affine* p) {
    bool zero = true;
    for (int i = 0; i < 12; i++) zero = zero && c[i] == 0;
    p->infinity = zero;
    if (zero) return true;
    if (!load_fq(c, &p->x) || !load_fq(c + 6, &p->y)) return false;
    const pr::Fq rhs = pr::add(pr::mul(pr::mul(p->x, p->x), p->x), pr::from_u64(4));
    return pr::eq(pr::mul(p->y, p->y), rhs);
}
bool load_g2(const uint64_t* c, pr::G2Affine* q) {
    bool zero = true;
    for (int i = 0; i < 24; i++) zero = zero && c[i] == 0;
    q->infinity = zero;
    if (zero) return true;
    if (!load_fq(c, &q->x.c0) || !load_fq(c + 6, &q->x.c1) || !load_fq(c + 12, &q->y.c0) || !load_fq(c + 18, &q->y.c1)) return false;
    const pr::Fq four = pr::from_u64(4);
    const pr::Fq2 b = {four, four};                                   // 4 (1 + u)
    return pr::eq(pr::mul(q->y, q->y), pr::add(pr::mul(pr::mul(q->x, q->x), q->x), b));
}
}  // namespace

// prod_i e(P_i, Q_i) == 1 ?   g1: n x 12 u64 (x, y), g2: n x 24 u64 (x.c0, x.c1, y.c0, y.c1): canonical little-endian limbs, all zero = the
// identity.  The pairing equation of MultilinearKZG::verify (multilinear_kzg.rs:90-116) in the form  e(C - v g1, -g2) prod_i e(pi_i, tau_i g2 - z_i g2) == 1.
// Points must lie on their curves (ZKSC_ERR_SHAPE otherwise); subgroup membership is the caller's business, as with ark-ec's unchecked types.
extern "C" int zksc_pairing_check(const uint64_t* g1, const uint64_t* g2, uint32_t n, int* is_one) {
    if (!is_one || (n && (!g1 || !g2))) return ZKSC_ERR_SHAPE;
    pr::Fq12 f = pr::kOne12;
    for (uint32_t i = 0; i < n; i++) {
        pr::G1Affine p;
        pr::G2Affine q;
        if (!load_g1(g1 + 12 * (size_t)i, &p) || !load_g2(g2 + 24 * (size_t)i, &q)) return ZKSC_ERR_SHAPE;
        f = pr::mul(f, pr::miller_loop(p, q));
    }
    *is_one = pr::eq(pr::final_exponentiation(f), pr::kOne12) ? 1 : 0;
    return ZKSC_OK;
}
// e(P, Q) itself: 12 Fq coefficients, canonical limbs, in the order ((c0.c0.c0, c0.c0.c1), (c0.c1.*), (c0.c2.*), (c1.c0.*), ...) of the tower above
extern "C" int zksc_pairing(const uint64_t* g1, const uint64_t* g2, uint64_t* out) {
    if (!g1 || !g2 || !out) return ZKSC_ERR_SHAPE;
    pr::G1Affine p;
    pr::G2Affine q;
    if (!load_g1(g1, &p) || !load_g2(g2, &q)) return ZKSC_ERR_SHAPE;
    const pr::Fq12 e = pr::final_exponentiation(pr::miller_loop(p, q));
    const pr::Fq6* h[2] = {&e.c0, &e.c1};
    for (int a = 0; a < 2; a++) {
        const pr::Fq2* t[3] = {&h[a]->c0, &h[a]->c1, &h[a]->c2};
        for (int b = 0; b < 3; b++) {
            pr::to_canonical(t[b]->c0, out + ((a * 3 + b) * 2) * 6);
            pr::to_canonical(t[b]->c1, out + ((a * 3 + b) * 2 + 1) * 6);
        }
    }
    return ZKSC_OK;
}

// MultilinearKZG::verify (kzg/src/multilinear_kzg.rs:90-116; sum_pairing_results, kzg/src/utils.rs:42-61) on the host:
//   e(commitment - evaluation g1, g2) == sum_i e(proof_i, tau_i g2 - point_i g2),   checked as   e(C - v g1, -g2) prod_i e(pi_i, tau_i g2 - z_i g2) == 1.
// commitment, proofs: ark-ec G1Projective memory form (18 u64 each, as zksc_g1_msm / zksc_kzg_open return them); points, evaluation: Montgomery
// Fr elements; srs_g2: the n G2 powers tau_i g2 of the trusted setup (trusted_setup.rs:37-46), affine canonical limbs (24 u64 each).
// *ok = 1 / 0.  ZKSC_ERR_SHAPE for malformed points.  No context, no device work.
extern "C" int zksc_kzg_verify(const uint64_t* commitment, const uint64_t* points, const uint64_t* evaluation, const uint64_t* proofs, const uint64_t* srs_g2,
                               uint32_t n_vars, int* ok) {
    if (!commitment || !evaluation || !ok || (n_vars && (!points || !proofs || !srs_g2))) return ZKSC_ERR_SHAPE;
    pr::G1Affine c;
    if (!pr::g1_from_ark(commitment, &c)) return ZKSC_ERR_SHAPE;
    const pr::G1Affine g1 = pr::g1_generator();
    const pr::G2Affine g2 = pr::g2_generator();
    uint64_t k[4];
    host::to_canonical(load_h(evaluation), k);
    pr::Fq12 f = pr::miller_loop(pr::g1_add(c, pr::g1_neg(pr::g1_mul(k, g1))), pr::g2_neg(g2));
    for (uint32_t i = 0; i < n_vars; i++) {
        pr::G1Affine pi;
        pr::G2Affine tau;
        if (!pr::g1_from_ark(proofs + 18 * (size_t)i, &pi) || !load_g2(srs_g2 + 24 * (size_t)i, &tau)) return ZKSC_ERR_SHAPE;
        host::to_canonical(load_h(points + 4 * (size_t)i), k);
        f = pr::mul(f, pr::miller_loop(pi, pr::g2_add(tau, pr::g2_neg(pr::g2_mul(k, g2)))));
    }
    *ok = pr::eq(pr::final_exponentiation(f), pr::kOne12) ? 1 : 0;
    return ZKSC_OK;
}
