// GKR layer driver inside the library: GKRProtocol::prove (gkr/src/protocol.rs:21-113) and
// generate_layer_one_prove_sumcheck (gkr/src/utils.rs:12-57) as ONE C call.  Included by zksc.cu only (it uses that
// file's table handle, launch helpers and host transcript).
//
// Per layer the reference builds four 2^(2k)-entry tables and calls MultiComposedSumcheckProver::prove_partial on
//     [ add~(b,c) * (W(b) + W(c)),  mul~(b,c) * (W(b) W(c)) ]                            (protocol.rs:67-93)
// Here:  W is uploaded once per layer and stays in HBM for the outer sum / outer product (add_distinct / mul_distinct,
// evaluation_form.rs:28-52) and for the two evaluations W(b*), W(c*) (one launch, gkr_eval_small_kernel);  the wiring
// tables are never materialised densely: folding the 0/1 table of 2^(3k-1) entries over its gate-label variables leaves
// sum_g eq(r, g) at entry (in0(g), in1(g)) -- at most one value per gate, formed on the host in exact field arithmetic
// (identical canonical values) and scattered into a zeroed table;  the layer sumcheck is zksc_prove (CUDA round kernels,
// its own transcript);  the outer GKR transcript stays on the host.
#pragma once

namespace zksc {

// W(point_j), j = blockIdx.x, for tables of at most 2^kGkrSmallLog entries: the first fold reads W from HBM, the
// remaining ones run in shared memory (Multilinear::evaluation = successive variable-0 folds, evaluation_form.rs:162-175).
constexpr int kGkrSmallLog = 11;
struct GkrPoints {
    Fr pt[2][kGkrSmallLog];    // the coordinates travel as kernel parameters: no host->device copy in front of the launch
};
__global__ void __launch_bounds__(256) gkr_eval_small_kernel(const Fr* w, unsigned int k, const __grid_constant__ GkrPoints points, Fr* out) {
    __shared__ Fr buf[1 << (kGkrSmallLog - 1)];
    const Fr* pt = points.pt[blockIdx.x];
    unsigned int half = 1u << (k - 1);
    {
        const Fr r = pt[0];
        for (unsigned int o = threadIdx.x; o < half; o += blockDim.x) buf[o] = fr_fold(ld256(w + o), ld256(w + o + half), r);
    }
    for (unsigned int j = 1; j < k; j++) {
        __syncthreads();
        half >>= 1;
        const Fr r = pt[j];
        // in place: entry o is read by the thread that writes it and by nobody else (the other operand is o + half >= half)
        for (unsigned int o = threadIdx.x; o < half; o += blockDim.x) {
            const Fr v = fr_fold(buf[o], buf[o + half], r);
            buf[o] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) st256(out + blockIdx.x, buf[0]);
}

}  // namespace zksc

// evaluations of a device-resident multilinear table at n_points points (each of k coordinates, on the device too)
static int gkr_eval_device(zksc_ctx* ctx, const Fr* d_w, uint32_t k, const uint64_t* h_points, uint32_t n_points, Fr* d_out, uint64_t* h_out) {
    if (k >= 1 && k <= (uint32_t)kGkrSmallLog && n_points <= 2) {
        GkrPoints pts;
        for (uint32_t q = 0; q < n_points; q++) memcpy(pts.pt[q], h_points + (size_t)q * k * 4, (size_t)k * 32);
        gkr_eval_small_kernel<<<n_points, 256, 0, ctx->stream>>>(d_w, k, pts, d_out);
        ctx->launches++;
        CK(cudaGetLastError());
    } else {
        // general size: one out-of-place variable-0 fold per point into scratch, then in-place folds (fold_kernel)
        const uint64_t n = 1ull << k;
        for (uint32_t q = 0; q < n_points; q++) {
            if (k == 0) { CK(cudaMemcpyAsync(d_out + q, d_w, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream)); continue; }
            DevBuf scratch(ctx);
            CK(dev_alloc(ctx, (void**)&scratch.p, n / 2 * sizeof(Fr)));
            uint64_t cur = n;
            for (uint32_t j = 0; j < k; j++) {
                FoldArgs a;
                a.in = j == 0 ? d_w : scratch.p; a.out = scratch.p;
                a.in_tab_stride = a.in_proof_stride = a.out_tab_stride = a.out_proof_stride = 0;
                a.n_out = cur / 2; a.s = cur / 2; a.n_tabs = 1;
                memcpy(a.chal[0].l, h_points + ((size_t)q * k + j) * 4, 32);
                fold_kernel<<<dim3(grid_for(ctx, a.n_out, 256, 8), 1), 256, 0, ctx->stream>>>(a);
                ctx->launches++;
                cur /= 2;
            }
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(d_out + q, scratch.p, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    CK(cudaMemcpyAsync(h_out, d_out, (size_t)n_points * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZKSC_OK;
}

// eq(r, a) for a = 0 .. 2^len - 1, a's most significant bit paired with r[0]: what partial_evaluations(r, [0; len])
// (evaluation_form.rs:143-159) leaves of the indicator of label a.
static std::vector<FrH> gkr_eq_vector(const std::vector<FrH>& r) {
    std::vector<FrH> v(1, host::kOne);
    for (const FrH& x : r) {
        const FrH nx = host::sub(host::kOne, x);
        std::vector<FrH> nv(v.size() * 2);
        for (size_t i = 0; i < v.size(); i++) { nv[2 * i] = host::mul(v[i], nx); nv[2 * i + 1] = host::mul(v[i], x); }
        v.swap(nv);
    }
    return v;
}

// Gate arrays and the two CSR orders (gates grouped by first / by second input) of one circuit layer, in HBM; gkr_linear.cuh has the
// layer sumcheck that works from them in two phases on tables of 2^k entries instead of 2^(2k).
struct GkrLayerDev {
    unsigned char* type = nullptr;                               // [gates]
    unsigned int *in0 = nullptr, *in1 = nullptr;                 // [gates]
    unsigned int *row0 = nullptr, *gate0 = nullptr;              // gates grouped by first input: [wires + 1], [gates]
    unsigned int *row1 = nullptr, *gate1 = nullptr;              // ... by second input
};
static int gkr_two_phase_layer(zksc_ctx* ctx, const GkrLayerDev& l, const Fr* d_w, uint32_t k, unsigned long long ng, const std::vector<FrH>& r_b,
                               const std::vector<FrH>* r_c, const FrH& alpha, const FrH& beta, Fr* wgt, Fr* eq_half, size_t hsz, zksc_tables* t, const FrH& claimed,
                               uint32_t stride, uint64_t* msgs, uint32_t* lens, uint64_t* chal, FrH* wu_out, FrH* wv_out, double* prof);

extern "C" uint64_t zksc_gkr_total_rounds(uint32_t n_layers) { return (uint64_t)n_layers * (n_layers + 1); }

extern "C" int zksc_gkr_prove(zksc_ctx* ctx, uint32_t n_layers, const uint32_t* n_gates, const uint8_t* gate_type, const uint32_t* gate_in0,
                              const uint32_t* gate_in1, const uint64_t* const* layer_values, const uint64_t* value_len, uint64_t* w0, uint64_t* sums,
                              uint64_t* wb_s, uint64_t* wc_s, uint64_t* round_msgs, uint32_t* round_len, uint64_t* challenges) {
    if (!ctx) return ZKSC_ERR_STATE;
    if (!ctx->kids.empty()) ctx = ctx->kids[0]->host_reduce ? nullptr : ctx;
    if (!ctx) return ZKSC_ERR_UNSUPPORTED;      // GKR layer sumchecks are far below the size where sharding pays: use a single-device context
    if (!n_gates || !gate_type || !gate_in0 || !gate_in1 || !layer_values || !value_len || !w0 || !sums || !wb_s || !wc_s || !round_msgs || !round_len ||
        !challenges)
        FAIL(ZKSC_ERR_SHAPE, "NULL argument");
    if (n_layers < 1 || n_layers > 15) FAIL(ZKSC_ERR_SHAPE, "GKR: 1..15 layers (layer i has 2^i gates; its sumcheck runs over 2^(2i+2) entries)");
    if (ctx->n_ranks != 1) FAIL(ZKSC_ERR_UNSUPPORTED, "GKR layer sumchecks run on one GPU per proof (independent proofs are replicas); use an unsharded context");
    // shapes the reference's constructors enforce: Multilinear::new (power of two, evaluation_form.rs:12-26) and
    // ComposedMultilinear::new (equal arity, composed_multilinear.rs:12-18) with the label widths of circuit/src/utils.rs:1-34
    if (value_len[0] != 1) FAIL(ZKSC_ERR_SHAPE, "GKR: the output layer has one gate (w_0 = [out, 0], protocol.rs:31-34)");
    size_t gate_off = 0;
    for (uint32_t li = 0; li < n_layers; li++) {
        if (!layer_values[li] || !layer_values[li + 1]) FAIL(ZKSC_ERR_SHAPE, "NULL layer values");
        if (value_len[li] != (1ull << li) || n_gates[li] != value_len[li]) FAIL(ZKSC_ERR_SHAPE, "GKR: layer i must have 2^i gates and 2^i values");
        if (value_len[li + 1] != (2ull << li)) FAIL(ZKSC_ERR_SHAPE, "GKR: layer i + 1 must have 2^(i+1) values (Number of evaluations must be a power of 2; equal arity)");
        for (uint32_t g = 0; g < n_gates[li]; g++) {
            if (gate_type[gate_off + g] > 1) FAIL(ZKSC_ERR_SHAPE, "gate type must be 0 (Add) or 1 (Mul)");
            if ((gate_in0[gate_off + g] >> (li + 1)) || (gate_in1[gate_off + g] >> (li + 1))) FAIL(ZKSC_ERR_SHAPE, "gate input label does not fit its bit field");
        }
        gate_off += n_gates[li];
    }
    CK(cudaSetDevice(ctx->device));
    TRY(quiesce(ctx));

    host::FiatShamirTranscript transcript;
    // w_0 = [out, 0]; transcript.commit(w_0.to_bytes()); n_r; claimed = w_0(n_r)          protocol.rs:31-38
    const FrH out0 = load_h(layer_values[0]);
    store_h(w0, out0);
    store_h(w0 + 4, host::kZero);
    transcript.commit_field(out0);
    transcript.commit_field(host::kZero);
    std::vector<FrH> r_b = transcript.evaluate_n_challenge_into_field(1), r_c;
    FrH claimed = host::mul(out0, host::sub(host::kOne, r_b[0]));   // (1 - r) out + r * 0
    FrH alpha = host::kOne, beta = host::kZero;

    const uint32_t degs[2] = {2, 2};
    const uint32_t stride = zksc_msg_stride(ZKSC_PROTO_MULTI_PARTIAL, 2, degs);
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> idx_add, idx_mul;
    std::vector<FrH> val_add, val_mul;
    size_t round_off = 0;
    gate_off = 0;
    for (uint32_t li = 0; li < n_layers; li++) {
        const uint32_t k = li + 1, n = 2 * k;
        const uint64_t nw = 1ull << k;
        const auto p0 = std::chrono::steady_clock::now();
        if (k >= ctx->gkr_linear_min) {
            // Large layers: the two-phase form of gkr_linear.cuh -- 2k rounds on tables of 2^k entries built from the gate lists instead of four
            // 2^(2k)-entry tables; the same round polynomials, hence the same bytes (tests/test_gpu_gkr.py compares both forms on every
            // reference circuit).  Measured on Circuit::random(10): the two largest layers 760 + 400 us -> 300 + 280 us.
            const size_t ng = n_gates[li];
            const uint32_t* g_in0 = gate_in0 + gate_off;
            const uint32_t* g_in1 = gate_in1 + gate_off;
            // one staging buffer [W | row0 | gate0 | row1 | gate1 | in0 | in1 | type], one copy
            const size_t n_u32 = 2 * (nw + 1) + 4 * ng, bytes_total = nw * sizeof(Fr) + n_u32 * 4 + ng;
            const size_t hsz = (size_t)1 << ((k + 1) / 2), n_scratch = std::max<size_t>(ng, 2) + 6 * hsz;
            if (ctx->gkr_stage_cap * sizeof(uint64_t) < bytes_total + 64) {
                CK(cudaStreamSynchronize(ctx->stream));
                if (ctx->gkr_stage) cudaFreeHost(ctx->gkr_stage);
                ctx->gkr_stage = nullptr; ctx->gkr_stage_cap = 0;
                CK(cudaHostAlloc((void**)&ctx->gkr_stage, 2 * (bytes_total + 64), cudaHostAllocDefault));
                ctx->gkr_stage_cap = 2 * (bytes_total + 64) / sizeof(uint64_t);
            }
            uint8_t* stage = (uint8_t*)ctx->gkr_stage;       // free again: the previous layer ended with a stream synchronisation
            memcpy(stage, layer_values[li + 1], nw * sizeof(Fr));
            uint32_t* row0 = (uint32_t*)(stage + nw * sizeof(Fr));
            uint32_t *gate0 = row0 + nw + 1, *row1 = gate0 + ng, *gate1 = row1 + nw + 1, *s_in0 = gate1 + ng, *s_in1 = s_in0 + ng;
            uint8_t* s_type = (uint8_t*)(s_in1 + ng);
            memset(row0, 0, (nw + 1) * 4);
            memset(row1, 0, (nw + 1) * 4);
            for (size_t g = 0; g < ng; g++) { row0[g_in0[g] + 1]++; row1[g_in1[g] + 1]++; }
            for (size_t w = 0; w < nw; w++) { row0[w + 1] += row0[w]; row1[w + 1] += row1[w]; }
            {
                std::vector<uint32_t> q0(row0, row0 + nw), q1(row1, row1 + nw);
                for (size_t g = 0; g < ng; g++) { gate0[q0[g_in0[g]]++] = (uint32_t)g; gate1[q1[g_in1[g]]++] = (uint32_t)g; }
            }
            memcpy(s_in0, g_in0, ng * 4);
            memcpy(s_in1, g_in1, ng * 4);
            memcpy(s_type, gate_type + gate_off, ng);
            DevBuf dbuf(ctx), scratch(ctx);
            CK(dev_alloc(ctx, (void**)&dbuf.p, bytes_total + 64));
            CK(dev_alloc(ctx, (void**)&scratch.p, n_scratch * sizeof(Fr)));
            CK(cudaMemcpyAsync(dbuf.p, stage, bytes_total, cudaMemcpyHostToDevice, ctx->stream));
            GkrLayerDev l;
            uint8_t* d = (uint8_t*)dbuf.p + nw * sizeof(Fr);
            l.row0 = (unsigned int*)d; l.gate0 = l.row0 + nw + 1; l.row1 = l.gate0 + ng; l.gate1 = l.row1 + nw + 1; l.in0 = l.gate1 + ng; l.in1 = l.in0 + ng;
            l.type = (unsigned char*)(l.in1 + ng);
            zksc_tables* t = nullptr;
            TRY(tables_alloc(ctx, k, 1, 2, degs, &t));
            struct Guard { zksc_tables* t; ~Guard() { if (t) zksc_tables_free(t); } } guard{t};
            store_h(sums + 4 * li, claimed);
            uint64_t* msgs = round_msgs + round_off * stride * 4;
            uint32_t* lens = round_len + round_off;
            uint64_t* chal = challenges + round_off * 4;
            FrH wu, wv;
            TRY(gkr_two_phase_layer(ctx, l, dbuf.p, k, ng, r_b, li > 0 ? &r_c : nullptr, alpha, beta, scratch.p, scratch.p + std::max<size_t>(ng, 2), hsz, t, claimed, stride,
                                    msgs, lens, chal, &wu, &wv, nullptr));
            const auto p2 = std::chrono::steady_clock::now();
            size_t blen = 0;
            TRY(zksc_proof_to_bytes(ZKSC_PROTO_MULTI_PARTIAL, n, stride, msgs, lens, nullptr, &blen));
            bytes.resize(blen);
            TRY(zksc_proof_to_bytes(ZKSC_PROTO_MULTI_PARTIAL, n, stride, msgs, lens, bytes.data(), &blen));
            transcript.commit(bytes);
            r_b.clear(); r_c.clear();
            for (uint32_t j = 0; j < k; j++) { r_b.push_back(load_h(chal + 4 * j)); r_c.push_back(load_h(chal + 4 * (k + j))); }
            store_h(wb_s + 4 * li, wu);
            store_h(wc_s + 4 * li, wv);
            alpha = transcript.evaluate_challenge_into_field();
            beta = transcript.evaluate_challenge_into_field();
            claimed = host::add(host::mul(alpha, wu), host::mul(beta, wv));
            round_off += n;
            gate_off += n_gates[li];
            if (ctx->profile) {
                auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
                fprintf(stderr, "[zksc profile] gkr layer %2u (two-phase form): tables + prove %7.1f us  absorb %7.1f us\n", li, us(p0, p2), us(p2, std::chrono::steady_clock::now()));
            }
            continue;
        }
        // wiring entries: layer one add(n_r, b, c) unscaled (utils.rs:23-24); later alpha add(r_b,.,.) + beta add(r_c,.,.) (protocol.rs:86-88)
        idx_add.clear(); idx_mul.clear(); val_add.clear(); val_mul.clear();
        {
            std::vector<FrH> eq = gkr_eq_vector(r_b);
            if (li > 0) {
                const std::vector<FrH> eq_c = gkr_eq_vector(r_c);
                for (size_t i = 0; i < eq.size(); i++) eq[i] = host::add(host::mul(eq[i], alpha), host::mul(eq_c[i], beta));
            }
            if (eq.size() < n_gates[li]) FAIL(ZKSC_ERR_SHAPE, "challenge vector does not match the layer's gate-label bits");
            // several gates may share one (in0, in1) entry: accumulate per entry
            std::vector<std::pair<uint64_t, uint32_t>> order(n_gates[li]);
            for (uint32_t g = 0; g < n_gates[li]; g++)
                order[g] = {((uint64_t)gate_type[gate_off + g] << 62) | ((uint64_t)gate_in0[gate_off + g] << k) | gate_in1[gate_off + g], g};
            std::sort(order.begin(), order.end());
            for (size_t i = 0; i < order.size(); i++) {
                const bool is_mul = order[i].first >> 62;
                const uint64_t d = order[i].first & ((1ull << 62) - 1);
                std::vector<uint64_t>& ix = is_mul ? idx_mul : idx_add;
                std::vector<FrH>& vl = is_mul ? val_mul : val_add;
                if (i > 0 && order[i].first == order[i - 1].first) vl.back() = host::add(vl.back(), eq[order[i].second]);
                else { ix.push_back(d); vl.push_back(eq[order[i].second]); }
            }
        }
        zksc_tables* t = nullptr;
        TRY(tables_alloc(ctx, n, 1, 2, degs, &t));
        struct Guard { zksc_tables* t; ~Guard() { if (t) zksc_tables_free(t); } } guard{t};
        // one device buffer [W | sparse values | sparse indices | 2 results], filled by ONE copy from a pinned staging buffer
        const size_t n_sp = idx_add.size() + idx_mul.size();
        const size_t off_val = nw, off_idx = off_val + n_sp, off_res = off_idx + (n_sp + 3) / 4, n_buf = off_res + 2;
        DevBuf dbuf(ctx);
        CK(dev_alloc(ctx, (void**)&dbuf.p, n_buf * sizeof(Fr)));
        if (ctx->gkr_stage_cap < off_res * 4) {
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->gkr_stage) cudaFreeHost(ctx->gkr_stage);
            ctx->gkr_stage = nullptr; ctx->gkr_stage_cap = 0;
            CK(cudaHostAlloc((void**)&ctx->gkr_stage, off_res * 4 * 2 * sizeof(uint64_t), cudaHostAllocDefault));
            ctx->gkr_stage_cap = off_res * 4 * 2;
        }
        {
            uint64_t* stage = ctx->gkr_stage;       // free again: the previous layer ended with a stream synchronisation
            memcpy(stage, layer_values[li + 1], nw * sizeof(Fr));
            for (size_t i = 0; i < idx_add.size(); i++) store_h(&stage[4 * (off_val + i)], val_add[i]);
            for (size_t i = 0; i < idx_mul.size(); i++) store_h(&stage[4 * (off_val + idx_add.size() + i)], val_mul[i]);
            memcpy(&stage[4 * off_idx], idx_add.data(), idx_add.size() * 8);
            memcpy(&stage[4 * off_idx + idx_add.size()], idx_mul.data(), idx_mul.size() * 8);
            CK(cudaMemcpyAsync(dbuf.p, stage, off_res * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
        }
        const Fr* d_w = dbuf.p;
        Fr* tab = t->orig;
        const uint64_t N = t->n_local0;
        // W(b) + W(c), W(b) W(c)                                                             protocol.rs:80-81
        outer_fill_kernel<<<grid_for(ctx, N, 256, 8), 256, 0, ctx->stream>>>(0, d_w, d_w, nw, tab + 1 * N, N, 0, 1);
        outer_fill_kernel<<<grid_for(ctx, N, 256, 8), 256, 0, ctx->stream>>>(1, d_w, d_w, nw, tab + 3 * N, N, 0, 1);
        ctx->launches += 2;
        CK(cudaMemsetAsync(tab, 0, N * sizeof(Fr), ctx->stream));
        CK(cudaMemsetAsync(tab + 2 * N, 0, N * sizeof(Fr), ctx->stream));
        if (n_sp) {
            const Fr* d_val = dbuf.p + off_val;
            const unsigned long long* didx = (const unsigned long long*)(dbuf.p + off_idx);
            if (!idx_add.empty()) {
                scatter_kernel<<<(unsigned int)((idx_add.size() + 255) / 256), 256, 0, ctx->stream>>>(didx, d_val, idx_add.size(), tab, 0, 1);
                ctx->launches++;
            }
            if (!idx_mul.empty()) {
                scatter_kernel<<<(unsigned int)((idx_mul.size() + 255) / 256), 256, 0, ctx->stream>>>(didx + idx_add.size(), d_val + idx_add.size(), idx_mul.size(),
                                                                                                       tab + 2 * N, 0, 1);
                ctx->launches++;
            }
        }
        CK(cudaGetLastError());
        t->r0_valid = false;

        // prove_partial(&[add~ (W+W), mul~ (W W)], &claimed_sum); transcript.commit(&proof.to_bytes())   protocol.rs:93-98
        uint64_t sum_m[4];
        store_h(sum_m, claimed);
        store_h(sums + 4 * li, claimed);
        uint64_t* msgs = round_msgs + round_off * stride * 4;
        uint32_t* lens = round_len + round_off;
        uint64_t* chal = challenges + round_off * 4;
        const auto p1 = std::chrono::steady_clock::now();
        TRY(zksc_prove(t, ZKSC_PROTO_MULTI_PARTIAL, sum_m, msgs, lens, chal));
        const auto p2 = std::chrono::steady_clock::now();
        size_t blen = 0;
        TRY(zksc_proof_to_bytes(ZKSC_PROTO_MULTI_PARTIAL, n, stride, msgs, lens, nullptr, &blen));
        bytes.resize(blen);
        TRY(zksc_proof_to_bytes(ZKSC_PROTO_MULTI_PARTIAL, n, stride, msgs, lens, bytes.data(), &blen));
        transcript.commit(bytes);
        TRY(tail_stop(t));   // the resident rounds kernel (if any) has applied the last challenge; the stream is free again

        // (b, c) = challenges.split_at(len / 2); W(b), W(c)                                    protocol.rs:100-105
        r_b.clear(); r_c.clear();
        for (uint32_t j = 0; j < k; j++) { r_b.push_back(load_h(chal + 4 * j)); r_c.push_back(load_h(chal + 4 * (k + j))); }
        uint64_t ev[8];
        TRY(gkr_eval_device(ctx, d_w, k, chal, 2, dbuf.p + off_res, ev));
        memcpy(wb_s + 4 * li, ev, 32);
        memcpy(wc_s + 4 * li, ev + 4, 32);
        alpha = transcript.evaluate_challenge_into_field();                                      // protocol.rs:110-111
        beta = transcript.evaluate_challenge_into_field();
        claimed = host::add(host::mul(alpha, load_h(ev)), host::mul(beta, load_h(ev + 4)));      // :113
        round_off += n;
        gate_off += n_gates[li];
        if (ctx->profile) {
            auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
            fprintf(stderr, "[zksc profile] gkr layer %2u: tables %7.1f us  prove %7.1f us  absorb + W(b), W(c) %7.1f us\n", li, us(p0, p1), us(p1, p2),
                    us(p2, std::chrono::steady_clock::now()));
        }
    }
    return ZKSC_OK;
}
