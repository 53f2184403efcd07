// GKR for layered circuits of ANY layer widths, in time linear in the number of gates -- BASELINE.json config 4 read literally
// ("random layered add/mul circuit, width 2^20, depth 8"; SURVEY.md 8(f) next-1, reading (B)).  Included by zksc.cu only.
//
// The reference's prover (gkr/src/protocol.rs:21-113) runs, per layer, MultiComposedSumcheckProver::prove_partial on
//     [ add~(b,c) * (W(b) + W(c)),  mul~(b,c) * (W(b) W(c)) ]                              over the 2k variables (b, c)
// with four dense tables of 2^(2k) entries, k = log2(width of the layer below).  That form (zksc_gkr_prove, gkr_driver.cuh) stops at
// k = 10-11; a layer of width 2^20 would need 2^40 entries.  The round polynomials, however, do not need the dense tables: the k
// rounds that bind b sum c out, the k rounds that bind c see b already fixed (Libra's two phases).  With
//     wgt(g) = alpha eq(r_b, g) + beta eq(r_c, g)                                          (protocol.rs:86-88; layer one: eq(n_r, g))
//   phase 1 (b):  sum_c f(b, c) = W(b) [ sum_{add g: in0 = b} wgt(g) + sum_{mul g: in0 = b} wgt(g) W(in1 g) ] + sum_{add g: in0 = b} wgt(g) W(in1 g)
//                               = W(b) H1(b) + H2(b) * 1
//   phase 2 (c):  f(u, c) = A(c) (W(u) + W(c)) + [W(u) M(c)] W(c),     A / M (c) = sum_{add / mul g: in1 = c} wgt(g) eq(u, in0 g)
// -- each phase a two-product, degree-2 sumcheck over 2^k entries: the SAME round polynomials as the dense form (they are the same
// polynomials in the round's variable), absorbed by the same transcript in the same order, so the proof bytes are the dense
// prover's.  tests/test_gpu_gkr.py checks exactly that: byte-identical to zksc_gkr_prove and to the oracle on the reference's
// pyramid circuits, and to the oracle's dense prover on uniform-width circuits small enough for it.
//
// The circuit is preprocessed once (zksc_circuit_create): per layer the gates grouped by their first and by their second input
// (two CSR orders, built on the host by counting sort, resident in HBM).  Circuit::evaluation (circuit/src/circuit.rs:32-55) runs on
// the device too (zksc_circuit_evaluate), one launch per layer, and the layer values stay in HBM for the prover.
#pragma once

namespace zksc {

constexpr int kEqMaxVars = 30;
constexpr int kEqHalfMax = 15;         // eq(r, a) = hi[a >> kl] lo[a & (2^kl - 1)]: two tables of at most 2^15 entries, ONE product per label
struct EqPoint {
    Fr r[kEqMaxVars];          // r[0] pairs with the label's most significant bit (evaluation_form.rs:143-159: successive variable-0 folds)
    unsigned int k;
};
// the two half tables of eq(p, .): hi over the first k - kl coordinates (times `scale`), lo over the last kl = k / 2
__global__ void __launch_bounds__(256) gkr_eq_halves_kernel(const __grid_constant__ EqPoint p, const Fr scale, Fr* hi, Fr* lo) {
    const unsigned int kl = p.k / 2, kh = p.k - kl;
    const unsigned int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < (1u << kh)) {
        Fr v = scale;
        for (unsigned int j = 0; j < kh; j++) {
            const Fr x = p.r[j];
            v = fr_mul(v, ((a >> (kh - 1 - j)) & 1u) ? x : fr_sub(fr_one(), x));
        }
        st256(hi + a, v);
    }
    if (a < (1u << kl)) {
        Fr v = fr_one();
        for (unsigned int j = 0; j < kl; j++) {
            const Fr x = p.r[kh + j];
            v = fr_mul(v, ((a >> (kl - 1 - j)) & 1u) ? x : fr_sub(fr_one(), x));
        }
        st256(lo + a, v);
    }
}
struct EqHalves {
    const Fr *hi, *lo;
    unsigned int kl;
};
ZKSC_DEV Fr eq_from_halves(const EqHalves& e, unsigned int a) { return fr_mul(ld256(e.hi + (a >> e.kl)), ld256(e.lo + (a & ((1u << e.kl) - 1u)))); }
// wgt[g] = alpha eq(r_b, g) + beta eq(r_c, g)  (the scales sit in the hi tables; two == 0: the first term alone)
__global__ void __launch_bounds__(256) gkr_wgt_kernel(const EqHalves b, const EqHalves c, int two, Fr* out, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        Fr v = eq_from_halves(b, (unsigned int)g);
        if (two) v = fr_add(v, eq_from_halves(c, (unsigned int)g));
        st256(out + g, v);
    }
}
// Circuit::evaluation, one layer: out[g] = in[in0 g] (+ or *) in[in1 g]
__global__ void __launch_bounds__(256) gkr_layer_eval_kernel(const unsigned char* type, const unsigned int* in0, const unsigned int* in1, const Fr* in, Fr* out,
                                                             unsigned long long n_gates) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_gates; g += stride) {
        const Fr x = ld256(in + in0[g]), y = ld256(in + in1[g]);
        st256(out + g, type[g] ? fr_mul(x, y) : fr_add(x, y));
    }
}
// Phase 1 tables of one layer: [W, H1 | H2, 1].  row[b] .. row[b + 1] index the gates whose first input is wire b.
__global__ void __launch_bounds__(256) gkr_phase1_kernel(const unsigned int* row, const unsigned int* gate, const unsigned char* type, const unsigned int* in1,
                                                         const Fr* wgt, const Fr* w, Fr* tab, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) {
        Fr h1 = fr_zero(), h2 = fr_zero();
        for (unsigned int i = row[b]; i < row[b + 1]; i++) {
            const unsigned int g = gate[i];
            const Fr wg = ld256(wgt + g), t = fr_mul(wg, ld256(w + in1[g]));
            if (type[g]) h1 = fr_add(h1, t);
            else { h1 = fr_add(h1, wg); h2 = fr_add(h2, t); }
        }
        st256(tab + b, ld256(w + b));
        st256(tab + n + b, h1);
        st256(tab + 2 * n + b, h2);
        st256(tab + 3 * n + b, fr_one());
    }
}
// Phase 2 tables: [A, W(u) + W | W(u) M, W].  row[c] .. row[c + 1] index the gates whose second input is wire c.
__global__ void __launch_bounds__(256) gkr_phase2_kernel(const unsigned int* row, const unsigned int* gate, const unsigned char* type, const unsigned int* in0,
                                                         const Fr* wgt, const EqHalves eq_u, const Fr* w, const Fr wu, Fr* tab, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        Fr a = fr_zero(), m = fr_zero();
        for (unsigned int i = row[c]; i < row[c + 1]; i++) {
            const unsigned int g = gate[i];
            const Fr t = fr_mul(ld256(wgt + g), eq_from_halves(eq_u, in0[g]));
            if (type[g]) m = fr_add(m, t);
            else a = fr_add(a, t);
        }
        const Fr wc = ld256(w + c);
        st256(tab + c, a);
        st256(tab + n + c, fr_add(wu, wc));
        st256(tab + 2 * n + c, fr_mul(wu, m));
        st256(tab + 3 * n + c, wc);
    }
}

// Verifier side: per-block partial sums of  wgt(g) eq(b, in0 g) eq(c, in1 g)  over the add gates (out[2 blk]) and the mul gates (out[2 blk + 1])
// of a layer = add~(r, b, c), mul~(r, b, c) of gkr/src/protocol.rs:131-133, 164-171 without the dense wiring tables
__global__ void __launch_bounds__(256) gkr_wiring_eval_kernel(const unsigned char* type, const unsigned int* in0, const unsigned int* in1, const EqHalves wb,
                                                              const EqHalves wc, int two, const EqHalves eb, const EqHalves ec, Fr* out, unsigned long long n_gates) {
    __shared__ Fr s_a[256], s_m[256];
    Fr a = fr_zero(), m = fr_zero();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_gates; g += stride) {
        Fr w = eq_from_halves(wb, (unsigned int)g);
        if (two) w = fr_add(w, eq_from_halves(wc, (unsigned int)g));
        const Fr t = fr_mul(w, fr_mul(eq_from_halves(eb, in0[g]), eq_from_halves(ec, in1[g])));
        if (type[g]) m = fr_add(m, t);
        else a = fr_add(a, t);
    }
    s_a[threadIdx.x] = a;
    s_m[threadIdx.x] = m;
    for (unsigned int d = 128; d >= 1; d >>= 1) {
        __syncthreads();
        if (threadIdx.x < d) {
            s_a[threadIdx.x] = fr_add(s_a[threadIdx.x], s_a[threadIdx.x + d]);
            s_m[threadIdx.x] = fr_add(s_m[threadIdx.x], s_m[threadIdx.x + d]);
        }
    }
    if (threadIdx.x == 0) {
        st256(out + 2 * blockIdx.x, s_a[0]);
        st256(out + 2 * blockIdx.x + 1, s_m[0]);
    }
}

}  // namespace zksc

struct zksc_circuit {
    zksc_ctx* ctx = nullptr;
    uint32_t n_layers = 0;
    std::vector<uint32_t> lw;                 // [n_layers + 1] log2 of the layer widths; layer n_layers = the inputs
    typedef GkrLayerDev Layer;                // (gkr_driver.cuh: gate arrays and the two CSR orders of a layer, in HBM)
    std::vector<Layer> layers;
    std::vector<Fr*> values;                  // [n_layers + 1] layer values in HBM (zksc_circuit_evaluate)
    bool evaluated = false;
    Fr *wgt = nullptr, *eq_half = nullptr;    // scratch: weights of the widest layer; six half tables of eq (2^15 entries each)
    Fr* bytes = nullptr;                      // scratch: the output layer as big-endian bytes
    uint8_t* h_stage = nullptr;               // pinned: [bytes | Montgomery values] of the output layer on their way to the host
    std::map<uint32_t, zksc_tables*> handles; // prover table handles by number of variables
};

extern "C" int zksc_circuit_free(zksc_circuit* c) {
    if (!c) return ZKSC_OK;
    zksc_ctx* ctx = c->ctx;
    cudaSetDevice(ctx->device);
    for (auto& h : c->handles) zksc_tables_free(h.second);
    cudaStreamSynchronize(ctx->stream);
    for (auto& l : c->layers) { cudaFree(l.type); cudaFree(l.in0); cudaFree(l.in1); cudaFree(l.row0); cudaFree(l.gate0); cudaFree(l.row1); cudaFree(l.gate1); }
    for (Fr* v : c->values) cudaFree(v);
    cudaFree(c->wgt); cudaFree(c->eq_half); cudaFree(c->bytes);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    delete c;
    return ZKSC_OK;
}

extern "C" int zksc_circuit_create(zksc_ctx* ctx, uint32_t n_layers, const uint32_t* log_width, const uint8_t* gate_type, const uint32_t* gate_in0,
                                   const uint32_t* gate_in1, zksc_circuit** out) {
    if (!ctx || !out) return ZKSC_ERR_STATE;
    *out = nullptr;
    if (!ctx->kids.empty() || ctx->n_ranks != 1) FAIL(ZKSC_ERR_UNSUPPORTED, "layered circuits run on one GPU per proof; use an unsharded single-device context");
    if (!log_width || !gate_type || !gate_in0 || !gate_in1) FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    if (n_layers < 1 || n_layers > 64) FAIL(ZKSC_ERR_SHAPE, "layered circuit: 1..64 layers");
    for (uint32_t i = 0; i <= n_layers; i++)
        if (log_width[i] > 28 || (i > 0 && log_width[i] < 1)) FAIL(ZKSC_ERR_SHAPE, "layer widths are 2^0..2^28 (2^1 at least below the output layer)");
    static_assert(2 * zksc::kEqHalfMax >= 28, "eq half tables too small for the widest layer");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    zksc_circuit* c = new zksc_circuit();
    c->ctx = ctx; c->n_layers = n_layers;
    c->lw.assign(log_width, log_width + n_layers + 1);
    c->layers.resize(n_layers);
    c->values.assign(n_layers + 1, nullptr);
    struct Guard { zksc_circuit* c; ~Guard() { if (c) zksc_circuit_free(c); } } guard{c};
    size_t off = 0;
    uint32_t widest = 0;
    for (uint32_t li = 0; li < n_layers; li++) {
        const size_t ng = (size_t)1 << log_width[li], nw = (size_t)1 << log_width[li + 1];
        widest = std::max(widest, std::max(log_width[li], log_width[li + 1]));
        std::vector<unsigned int> row0(nw + 1, 0), row1(nw + 1, 0), g0(ng), g1(ng);
        for (size_t g = 0; g < ng; g++) {
            if (gate_type[off + g] > 1) FAIL(ZKSC_ERR_SHAPE, "gate type must be 0 (Add) or 1 (Mul)");
            if (gate_in0[off + g] >= nw || gate_in1[off + g] >= nw) FAIL(ZKSC_ERR_SHAPE, "gate input label does not fit the layer below");
            row0[gate_in0[off + g] + 1]++;
            row1[gate_in1[off + g] + 1]++;
        }
        for (size_t w = 0; w < nw; w++) { row0[w + 1] += row0[w]; row1[w + 1] += row1[w]; }
        {
            std::vector<unsigned int> p0(row0.begin(), row0.end() - 1), p1(row1.begin(), row1.end() - 1);
            for (size_t g = 0; g < ng; g++) { g0[p0[gate_in0[off + g]]++] = (unsigned int)g; g1[p1[gate_in1[off + g]]++] = (unsigned int)g; }   // gate order kept inside a group
        }
        zksc_circuit::Layer& L = c->layers[li];
        CK(cudaMalloc(&L.type, ng)); CK(cudaMalloc(&L.in0, ng * 4)); CK(cudaMalloc(&L.in1, ng * 4));
        CK(cudaMalloc(&L.row0, (nw + 1) * 4)); CK(cudaMalloc(&L.gate0, ng * 4)); CK(cudaMalloc(&L.row1, (nw + 1) * 4)); CK(cudaMalloc(&L.gate1, ng * 4));
        CK(cudaMemcpy(L.type, gate_type + off, ng, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.in0, gate_in0 + off, ng * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.in1, gate_in1 + off, ng * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.row0, row0.data(), (nw + 1) * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.gate0, g0.data(), ng * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.row1, row1.data(), (nw + 1) * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.gate1, g1.data(), ng * 4, cudaMemcpyHostToDevice));
        off += ng;
    }
    for (uint32_t i = 0; i <= n_layers; i++) CK(cudaMalloc(&c->values[i], sizeof(Fr) << log_width[i]));
    CK(cudaMalloc(&c->wgt, sizeof(Fr) << std::max(widest, 1u)));
    CK(cudaMalloc(&c->eq_half, 6 * (sizeof(Fr) << zksc::kEqHalfMax)));
    CK(cudaMalloc(&c->bytes, (sizeof(Fr) << std::max(log_width[0], 1u)) + 2 * sizeof(Fr)));
    CK(cudaHostAlloc((void**)&c->h_stage, (size_t)2 * (sizeof(Fr) << std::max(log_width[0], 1u)), cudaHostAllocDefault));
    guard.c = nullptr;
    *out = c;
    return ZKSC_OK;
}

// Circuit::evaluation (circuit/src/circuit.rs:32-55) on the device; the layer values stay in HBM for zksc_gkr_prove_linear
extern "C" int zksc_circuit_evaluate(zksc_circuit* c, const uint64_t* inputs, uint64_t* outputs) {
    if (!c || !inputs) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = c->ctx;
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    const uint32_t L = c->n_layers;
    CK(cudaMemcpyAsync(c->values[L], inputs, sizeof(Fr) << c->lw[L], cudaMemcpyHostToDevice, ctx->stream));
    for (uint32_t li = L; li-- > 0;) {
        const unsigned long long ng = 1ull << c->lw[li];
        const zksc_circuit::Layer& l = c->layers[li];
        zksc::gkr_layer_eval_kernel<<<grid_for(ctx, ng, 256, 8), 256, 0, ctx->stream>>>(l.type, l.in0, l.in1, c->values[li + 1], c->values[li], ng);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    if (outputs) CK(cudaMemcpyAsync(outputs, c->values[0], sizeof(Fr) << c->lw[0], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    c->evaluated = true;
    return ZKSC_OK;
}

extern "C" int zksc_circuit_layer_values(zksc_circuit* c, uint32_t layer, uint64_t* out) {
    if (!c || !out) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = c->ctx;
    if (layer > c->n_layers) FAIL(ZKSC_ERR_SHAPE, "no such layer");
    if (!c->evaluated) FAIL(ZKSC_ERR_STATE, "zksc_circuit_evaluate first");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    CK(cudaMemcpyAsync(out, c->values[layer], sizeof(Fr) << c->lw[layer], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

// rounds of all layer sumchecks: 2 log2(width of the layer below) per layer
extern "C" uint64_t zksc_circuit_total_rounds(const zksc_circuit* c) {
    uint64_t n = 0;
    if (c) for (uint32_t li = 0; li < c->n_layers; li++) n += 2ull * c->lw[li + 1];
    return n;
}

static zksc::EqPoint gkr_eq_point(const std::vector<FrH>& r) {
    zksc::EqPoint p;
    memset(&p, 0, sizeof(p));
    p.k = (unsigned int)r.size();
    for (size_t j = 0; j < r.size(); j++) store_h((uint64_t*)p.r[j].l, r[j]);
    return p;
}

// GKRProtocol::prove (gkr/src/protocol.rs:21-113) for the layered circuit `c` on the values of the latest zksc_circuit_evaluate.  Outputs as
// zksc_gkr_prove's: w0 (max(2, width of layer 0) elements: the output layer, a single output padded with 0 as protocol.rs:31-34 does), and per layer
// the claimed sum, W(b*), W(c*), and the rounds of its sumcheck (zksc_circuit_total_rounds in all; message stride zksc_msg_stride(MULTI_PARTIAL, 2, {2,2})).
// One layer sumcheck in two phases (see the head of this file) on the handle `t` (k variables, two degree-2 products): 2k rounds on one
// transcript, written to msgs / lens / chal like zksc_prove's; W(u) and W(v) come back as the residuals of the two phases.
// r_c == nullptr: the output layer's single point (wgt = eq(r_b, .), utils.rs:23-24), else wgt = alpha eq(r_b, .) + beta eq(r_c, .).
// wgt: ng elements of scratch; eq_half: six half tables of hsz elements each (hsz >= 2^ceil(max(k, |r_b|) / 2)).
static int gkr_two_phase_layer(zksc_ctx* ctx, const GkrLayerDev& l, const Fr* d_w, uint32_t k, unsigned long long ng, const std::vector<FrH>& r_b,
                               const std::vector<FrH>* r_c, const FrH& alpha, const FrH& beta, Fr* wgt, Fr* eq_half, size_t hsz, zksc_tables* t, const FrH& claimed,
                               uint32_t stride, uint64_t* msgs, uint32_t* lens, uint64_t* chal, FrH* wu_out, FrH* wv_out, double* prof) {
    const unsigned long long nw = 1ull << k;
    const auto p0 = std::chrono::steady_clock::now();
    auto us_since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count(); };
    auto halves = [&](const std::vector<FrH>& r, const FrH& scale, int slot) {
        const zksc::EqPoint p = gkr_eq_point(r);
        Fr sc;
        store_h((uint64_t*)sc.l, scale);
        const unsigned int kl = p.k / 2, kh = p.k - kl;
        zksc::gkr_eq_halves_kernel<<<((1u << kh) + 255) / 256, 256, 0, ctx->stream>>>(p, sc, eq_half + (2 * slot) * hsz, eq_half + (2 * slot + 1) * hsz);
        ctx->launches++;
        return zksc::EqHalves{eq_half + (2 * slot) * hsz, eq_half + (2 * slot + 1) * hsz, kl};
    };
    {
        const zksc::EqHalves hb = halves(r_b, alpha, 0), hc = r_c ? halves(*r_c, beta, 1) : hb;
        zksc::gkr_wgt_kernel<<<grid_for(ctx, ng, 256, 8), 256, 0, ctx->stream>>>(hb, hc, r_c ? 1 : 0, wgt, ng);
    }
    // ---- phase 1: the k rounds that bind b
    TRY(zksc_tables_reset(t));
    zksc::gkr_phase1_kernel<<<grid_for(ctx, nw, 256, 8), 256, 0, ctx->stream>>>(l.row0, l.gate0, l.type, l.in1, wgt, d_w, t->orig, nw);
    ctx->launches += 2;
    CK(cudaGetLastError());
    t->r0_valid = false;
    std::vector<host::FiatShamirTranscript> tr(1);
    tr[0].commit_field(claimed);                                          // multi_composed_sumcheck.rs:70
    // prove_run indexes its outputs by (proof * n + round) with the handle's own round count: this proof has 2k rounds, so the two runs
    // are given the layer's slots directly
    TRY(prove_run(t, ZKSC_PROTO_MULTI_PARTIAL, tr, k, 0, stride, msgs, lens, chal));
    if (prof) prof[0] = us_since(p0);
    uint64_t resid[16];
    TRY(zksc_residual(t, resid));                                         // W(u), H1(u), H2(u), 1
    if (prof) prof[1] = us_since(p0);
    const FrH wu = load_h(resid);
    std::vector<FrH> u(k);
    for (uint32_t j = 0; j < k; j++) u[j] = load_h(chal + 4 * j);
    // ---- phase 2: the k rounds that bind c
    {
        Fr wu_m;
        store_h((uint64_t*)wu_m.l, wu);
        const zksc::EqHalves hu = halves(u, host::kOne, 2);
        TRY(zksc_tables_reset(t));
        zksc::gkr_phase2_kernel<<<grid_for(ctx, nw, 256, 8), 256, 0, ctx->stream>>>(l.row1, l.gate1, l.type, l.in0, wgt, hu, d_w, wu_m, t->orig, nw);
        ctx->launches++;
        CK(cudaGetLastError());
        t->r0_valid = false;
    }
    TRY(prove_run(t, ZKSC_PROTO_MULTI_PARTIAL, tr, k, 0, stride, msgs + (size_t)k * stride * 4, lens + k, chal + (size_t)k * 4));
    if (prof) prof[2] = us_since(p0);
    TRY(zksc_residual(t, resid));                                         // A(v), W(u) + W(v), W(u) M(v), W(v)
    if (prof) prof[3] = us_since(p0);
    *wu_out = wu;
    *wv_out = load_h(resid + 12);
    return ZKSC_OK;
}

extern "C" int zksc_gkr_prove_linear(zksc_circuit* c, uint64_t* w0, uint64_t* sums, uint64_t* wb_s, uint64_t* wc_s, uint64_t* round_msgs, uint32_t* round_len,
                                     uint64_t* challenges) {
    if (!c) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = c->ctx;
    if (!w0 || !sums || !wb_s || !wc_s || !round_msgs || !round_len || !challenges) FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    if (!c->evaluated) FAIL(ZKSC_ERR_STATE, "zksc_circuit_evaluate first");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    const uint32_t L = c->n_layers;

    host::FiatShamirTranscript transcript;
    // w_0 = the output layer (one output: [out, 0]); transcript.commit(w_0.to_bytes()); n_r; claimed = w_0(n_r)          protocol.rs:31-38
    const auto q0 = std::chrono::steady_clock::now();
    auto us_since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count(); };
    // (to_bytes on the device, Multilinear::to_bytes evaluation_form.rs:54-62; the SHA-256 of the bytes and the challenges on the host;
    //  w_0(n_r) by successive folds on the device)
    const uint32_t k0 = std::max(c->lw[0], 1u);
    const size_t n0 = (size_t)1 << k0;
    Fr* d_w0 = c->bytes + n0;          // scratch behind the bytes: [out, 0] when the layer has one gate
    std::vector<FrH> r_b, r_c;
    FrH claimed;
    {
        const Fr* w0_src = c->values[0];
        if (c->lw[0] == 0) {
            CK(cudaMemsetAsync(d_w0, 0, 2 * sizeof(Fr), ctx->stream));
            CK(cudaMemcpyAsync(d_w0, c->values[0], sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
            w0_src = d_w0;
        }
        to_bytes_kernel<<<grid_for(ctx, n0, 256, 8), 256, 0, ctx->stream>>>(w0_src, c->bytes, n0);
        ctx->launches++;
        CK(cudaGetLastError());
        // through pinned memory (a pageable destination makes these two copies 20 ms instead of 1.2 for a layer of 2^20 values)
        CK(cudaMemcpyAsync(c->h_stage, c->bytes, n0 * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(c->h_stage + n0 * 32, w0_src, n0 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const double t_copy = us_since(q0);
        transcript.commit(c->h_stage, n0 * 32);
        r_b = transcript.evaluate_n_challenge_into_field(k0);
        memcpy(w0, c->h_stage + n0 * 32, n0 * sizeof(Fr));
        if (ctx->profile) fprintf(stderr, "[zksc profile] gkr linear: output layer to_bytes + copies %8.1f us, sha-256 + challenges %8.1f us\n", t_copy, us_since(q0) - t_copy);
        std::vector<uint64_t> pts((size_t)4 * k0);
        for (uint32_t j = 0; j < k0; j++) store_h(&pts[4 * j], r_b[j]);
        uint64_t ev0[4];
        DevBuf res(ctx);
        CK(dev_alloc(ctx, (void**)&res.p, sizeof(Fr)));
        TRY(gkr_eval_device(ctx, w0_src, k0, pts.data(), 1, res.p, ev0));
        claimed = load_h(ev0);
        if (ctx->profile) fprintf(stderr, "[zksc profile] gkr linear: output layer in all %8.1f us\n", us_since(q0));
    }
    FrH alpha = host::kOne, beta = host::kZero;

    const uint32_t degs[2] = {2, 2};
    const uint32_t stride = zksc_msg_stride(ZKSC_PROTO_MULTI_PARTIAL, 2, degs);
    memset(round_msgs, 0, (size_t)zksc_circuit_total_rounds(c) * stride * 32);
    ctx->round_us.assign((size_t)zksc_circuit_total_rounds(c), 0.0);
    std::vector<uint8_t> bytes;
    size_t round_off = 0;
    for (uint32_t li = 0; li < L; li++) {
        const uint32_t ka = std::max(c->lw[li], 1u), k = c->lw[li + 1], n = 2 * k;
        const unsigned long long ng = 1ull << c->lw[li];
        const zksc_circuit::Layer& l = c->layers[li];
        if (r_b.size() != ka) FAIL(ZKSC_ERR_STATE, "challenge vector does not match the layer's gate-label bits");
        const auto p0 = std::chrono::steady_clock::now();
        zksc_tables*& t = c->handles[k];
        if (!t) TRY(tables_alloc(ctx, k, 1, 2, degs, &t));
        store_h(sums + 4 * li, claimed);
        uint64_t* msgs = round_msgs + round_off * stride * 4;
        uint32_t* lens = round_len + round_off;
        uint64_t* chal = challenges + round_off * 4;
        FrH wu, wv;
        double prof[4] = {0, 0, 0, 0};
        TRY(gkr_two_phase_layer(ctx, l, c->values[li + 1], k, ng, r_b, li > 0 ? &r_c : nullptr, alpha, beta, c->wgt, c->eq_half, (size_t)1 << zksc::kEqHalfMax, t, claimed,
                                stride, msgs, lens, chal, &wu, &wv, prof));
        const double t_ph1 = prof[0], t_res1 = prof[1], t_ph2 = prof[2], t_res2 = prof[3];
        std::vector<FrH> u(k), v(k);
        for (uint32_t j = 0; j < k; j++) { u[j] = load_h(chal + 4 * j); v[j] = load_h(chal + 4 * (k + j)); }
        // transcript.commit(&proof.to_bytes()); W(b*), W(c*); alpha, beta                                    protocol.rs:93-113
        size_t blen = 0;
        TRY(zksc_proof_to_bytes(ZKSC_PROTO_MULTI_PARTIAL, n, stride, msgs, lens, nullptr, &blen));
        bytes.resize(blen);
        TRY(zksc_proof_to_bytes(ZKSC_PROTO_MULTI_PARTIAL, n, stride, msgs, lens, bytes.data(), &blen));
        transcript.commit(bytes);
        store_h(wb_s + 4 * li, wu);
        store_h(wc_s + 4 * li, wv);
        r_b = u; r_c = v;
        alpha = transcript.evaluate_challenge_into_field();
        beta = transcript.evaluate_challenge_into_field();
        claimed = host::add(host::mul(alpha, wu), host::mul(beta, wv));
        round_off += n;
        if (ctx->profile)
            fprintf(stderr, "[zksc profile] gkr linear layer %2u: tables + phase 1 %8.1f us, residual %7.1f, tables + phase 2 %8.1f, residual %7.1f, absorb %7.1f\n", li, t_ph1,
                    t_res1 - t_ph1, t_ph2 - t_res1, t_res2 - t_ph2, us_since(p0) - t_res2);
    }
    return ZKSC_OK;
}

// Verifier side of a layer (GKRProtocol::verify, gkr/src/protocol.rs:131-133 and :164-171):  out[0] = alpha add(r_b, b, c) + beta add(r_c, b, c),
// out[1] = the same for mul, evaluated from the gate lists on the device.  r_b, r_c: log2(width of `layer`) coordinates each (at least one; beta and
// r_c may be NULL: one point, as for the output layer); b, c: log2(width of layer + 1) coordinates each.  Montgomery elements.
extern "C" int zksc_circuit_wiring_eval(zksc_circuit* c, uint32_t layer, const uint64_t* r_b, const uint64_t* alpha, const uint64_t* r_c, const uint64_t* beta,
                                        const uint64_t* b, const uint64_t* cpt, uint64_t* out) {
    if (!c) return ZKSC_ERR_STATE;
    zksc_ctx* ctx = c->ctx;
    if (!r_b || !alpha || !b || !cpt || !out || (r_c && !beta)) FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    if (layer >= c->n_layers) FAIL(ZKSC_ERR_SHAPE, "no such layer");
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));
    const uint32_t ka = std::max(c->lw[layer], 1u), k = c->lw[layer + 1];
    const unsigned long long ng = 1ull << c->lw[layer];
    const zksc_circuit::Layer& l = c->layers[layer];
    const size_t hsz = (size_t)1 << zksc::kEqHalfMax;
    DevBuf tabs(ctx);
    const int blocks = grid_for(ctx, ng, 256, 4);
    CK(dev_alloc(ctx, (void**)&tabs.p, (8 * hsz + 2 * (size_t)blocks) * sizeof(Fr)));     // own half tables: a proof's may be in use by nobody, but keep them apart
    auto halves = [&](const uint64_t* pt, uint32_t kk, const FrH& scale, int slot) {
        std::vector<FrH> r(kk);
        for (uint32_t j = 0; j < kk; j++) r[j] = load_h(pt + 4 * j);
        const zksc::EqPoint p = gkr_eq_point(r);
        Fr sc;
        store_h((uint64_t*)sc.l, scale);
        const unsigned int kl = p.k / 2, kh = p.k - kl;
        zksc::gkr_eq_halves_kernel<<<((1u << kh) + 255) / 256, 256, 0, ctx->stream>>>(p, sc, tabs.p + (2 * slot) * hsz, tabs.p + (2 * slot + 1) * hsz);
        ctx->launches++;
        return zksc::EqHalves{tabs.p + (2 * slot) * hsz, tabs.p + (2 * slot + 1) * hsz, kl};
    };
    const zksc::EqHalves hb = halves(r_b, ka, load_h(alpha), 0), hc = r_c ? halves(r_c, ka, load_h(beta), 1) : hb;
    const zksc::EqHalves eb = halves(b, k, host::kOne, 2), ec = halves(cpt, k, host::kOne, 3);
    Fr* part = tabs.p + 8 * hsz;
    zksc::gkr_wiring_eval_kernel<<<blocks, 256, 0, ctx->stream>>>(l.type, l.in0, l.in1, hb, hc, r_c ? 1 : 0, eb, ec, part, ng);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<uint64_t> h((size_t)blocks * 8);
    CK(cudaMemcpyAsync(h.data(), part, (size_t)blocks * 2 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    FrH a = host::kZero, m = host::kZero;
    for (int i = 0; i < blocks; i++) { a = host::add(a, load_h(&h[8 * (size_t)i])); m = host::add(m, load_h(&h[8 * (size_t)i + 4])); }
    store_h(out, a);
    store_h(out + 4, m);
    return ZKSC_OK;
}
