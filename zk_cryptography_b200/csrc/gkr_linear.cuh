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
struct EqPoint {
    Fr r[kEqMaxVars];          // r[0] pairs with the label's most significant bit (evaluation_form.rs:143-159: successive variable-0 folds)
    unsigned int k;
};
ZKSC_DEV Fr eq_at(const EqPoint& p, unsigned long long a) {
    Fr v = fr_one();
    for (unsigned int j = 0; j < p.k; j++) {
        const Fr x = p.r[j];
        v = fr_mul(v, ((a >> (p.k - 1 - j)) & 1ull) ? x : fr_sub(fr_one(), x));
    }
    return v;
}
// out[a] = alpha eq(pb, a) + beta eq(pc, a)   (two == 0: eq(pb, a) alone)
__global__ void __launch_bounds__(256) gkr_eq_kernel(const __grid_constant__ EqPoint pb, const __grid_constant__ EqPoint pc, const Fr alpha, const Fr beta,
                                                     int two, Fr* out, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long a = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += stride) {
        Fr v = eq_at(pb, a);
        if (two) v = fr_add(fr_mul(v, alpha), fr_mul(eq_at(pc, a), beta));
        st256(out + a, v);
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
                                                         const Fr* wgt, const Fr* eq_u, const Fr* w, const Fr wu, Fr* tab, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        Fr a = fr_zero(), m = fr_zero();
        for (unsigned int i = row[c]; i < row[c + 1]; i++) {
            const unsigned int g = gate[i];
            const Fr t = fr_mul(ld256(wgt + g), ld256(eq_u + in0[g]));
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

}  // namespace zksc

struct zksc_circuit {
    zksc_ctx* ctx = nullptr;
    uint32_t n_layers = 0;
    std::vector<uint32_t> lw;                 // [n_layers + 1] log2 of the layer widths; layer n_layers = the inputs
    struct Layer {
        unsigned char* type = nullptr;        // [gates]
        unsigned int *in0 = nullptr, *in1 = nullptr;                 // [gates]
        unsigned int *row0 = nullptr, *gate0 = nullptr;              // gates grouped by first input: [wires + 1], [gates]
        unsigned int *row1 = nullptr, *gate1 = nullptr;              // ... by second input
    };
    std::vector<Layer> layers;
    std::vector<Fr*> values;                  // [n_layers + 1] layer values in HBM (zksc_circuit_evaluate)
    bool evaluated = false;
    Fr *wgt = nullptr, *eq_u = nullptr;       // scratch: widest layer
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
    cudaFree(c->wgt); cudaFree(c->eq_u);
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
    CK(cudaMalloc(&c->eq_u, sizeof(Fr) << std::max(widest, 1u)));
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
    const uint32_t k0 = std::max(c->lw[0], 1u);
    std::vector<uint64_t> w0h((size_t)4 << k0, 0);
    CK(cudaMemcpyAsync(w0h.data(), c->values[0], sizeof(Fr) << c->lw[0], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(w0, w0h.data(), w0h.size() * 8);
    std::vector<FrH> cur((size_t)1 << k0);
    for (size_t i = 0; i < cur.size(); i++) { cur[i] = load_h(&w0h[4 * i]); transcript.commit_field(cur[i]); }
    std::vector<FrH> r_b = transcript.evaluate_n_challenge_into_field(k0), r_c;
    for (uint32_t j = 0; j < k0; j++) {        // Multilinear::evaluation: successive variable-0 folds (evaluation_form.rs:162-175)
        const size_t half = cur.size() / 2;
        std::vector<FrH> nx(half);
        for (size_t i = 0; i < half; i++) nx[i] = host::add(cur[i], host::mul(r_b[j], host::sub(cur[i + half], cur[i])));
        cur.swap(nx);
    }
    FrH claimed = cur[0];
    FrH alpha = host::kOne, beta = host::kZero;

    const uint32_t degs[2] = {2, 2};
    const uint32_t stride = zksc_msg_stride(ZKSC_PROTO_MULTI_PARTIAL, 2, degs);
    memset(round_msgs, 0, (size_t)zksc_circuit_total_rounds(c) * stride * 32);
    ctx->round_us.assign((size_t)zksc_circuit_total_rounds(c), 0.0);
    std::vector<uint8_t> bytes;
    std::vector<host::FiatShamirTranscript> tr(1);
    size_t round_off = 0;
    for (uint32_t li = 0; li < L; li++) {
        const uint32_t ka = std::max(c->lw[li], 1u), k = c->lw[li + 1], n = 2 * k;
        const unsigned long long ng = 1ull << c->lw[li], nw = 1ull << k;
        const zksc_circuit::Layer& l = c->layers[li];
        if (r_b.size() != ka) FAIL(ZKSC_ERR_STATE, "challenge vector does not match the layer's gate-label bits");
        zksc_tables*& t = c->handles[k];
        if (!t) TRY(tables_alloc(ctx, k, 1, 2, degs, &t));
        // wgt(g) = alpha eq(r_b, g) + beta eq(r_c, g); the output layer: eq(n_r, g) (utils.rs:23-24, protocol.rs:86-88)
        {
            const zksc::EqPoint pb = gkr_eq_point(r_b), pc = gkr_eq_point(li > 0 ? r_c : r_b);
            Fr a_m, b_m;
            store_h((uint64_t*)a_m.l, alpha); store_h((uint64_t*)b_m.l, beta);
            zksc::gkr_eq_kernel<<<grid_for(ctx, ng, 256, 8), 256, 0, ctx->stream>>>(pb, pc, a_m, b_m, li > 0 ? 1 : 0, c->wgt, ng);
        }
        // ---- phase 1: the k rounds that bind b
        TRY(zksc_tables_reset(t));
        zksc::gkr_phase1_kernel<<<grid_for(ctx, nw, 256, 8), 256, 0, ctx->stream>>>(l.row0, l.gate0, l.type, l.in1, c->wgt, c->values[li + 1], t->orig, nw);
        ctx->launches += 2;
        CK(cudaGetLastError());
        t->r0_valid = false;
        store_h(sums + 4 * li, claimed);
        uint64_t* msgs = round_msgs + round_off * stride * 4;
        uint32_t* lens = round_len + round_off;
        uint64_t* chal = challenges + round_off * 4;
        tr[0] = host::FiatShamirTranscript();
        tr[0].commit_field(claimed);                                          // multi_composed_sumcheck.rs:70
        // prove_run indexes its outputs by (proof * n + round) with the handle's own round count: this proof has n = 2k rounds, so the two runs
        // are given the layer's slots directly
        TRY(prove_run(t, ZKSC_PROTO_MULTI_PARTIAL, tr, k, 0, stride, msgs, lens, chal));
        uint64_t resid[16];
        TRY(zksc_residual(t, resid));                                         // W(u), H1(u), H2(u), 1
        const FrH wu = load_h(resid);
        std::vector<FrH> u(k), v(k);
        for (uint32_t j = 0; j < k; j++) u[j] = load_h(chal + 4 * j);
        // ---- phase 2: the k rounds that bind c
        {
            const zksc::EqPoint pu = gkr_eq_point(u);
            Fr one_m, wu_m;
            store_h((uint64_t*)one_m.l, host::kOne); store_h((uint64_t*)wu_m.l, wu);
            zksc::gkr_eq_kernel<<<grid_for(ctx, nw, 256, 8), 256, 0, ctx->stream>>>(pu, pu, one_m, one_m, 0, c->eq_u, nw);
            TRY(zksc_tables_reset(t));
            zksc::gkr_phase2_kernel<<<grid_for(ctx, nw, 256, 8), 256, 0, ctx->stream>>>(l.row1, l.gate1, l.type, l.in0, c->wgt, c->eq_u, c->values[li + 1], wu_m,
                                                                                         t->orig, nw);
            ctx->launches += 2;
            CK(cudaGetLastError());
            t->r0_valid = false;
        }
        TRY(prove_run(t, ZKSC_PROTO_MULTI_PARTIAL, tr, k, 0, stride, msgs + (size_t)k * stride * 4, lens + k, chal + (size_t)k * 4));
        TRY(zksc_residual(t, resid));                                         // A(v), W(u) + W(v), W(u) M(v), W(v)
        const FrH wv = load_h(resid + 12);
        for (uint32_t j = 0; j < k; j++) v[j] = load_h(chal + 4 * (k + j));
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
    }
    return ZKSC_OK;
}
