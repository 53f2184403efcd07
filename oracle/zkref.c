/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's sumcheck path.
 *
 * Used by tests/ (parity checker), __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg.  The product (zk_cryptography_b200/) never links or calls this file.
 *
 * PARITY STATUS: "parity unpinned" at the protocol level -- the reference is pure Rust and cannot be
 * built here (no cargo/rustc, crates not vendored); its field arithmetic lives in the third-party
 * crates ark-ff 0.4.2 / ark-test-curves 0.4.2 (bls12_381::Fr, Cargo.toml:21,32), its hash in sha2
 * 0.10.8 (Cargo.toml:23); no reference test pins a challenge, round polynomial or proof bytes.
 * Pinned: all primitive known-answer tests of the reference (tests/test_oracle_kat.py) and byte-for-byte
 * agreement with the independent big-int model oracle/pymodel.py (tests/test_oracle_cross.py).
 *
 * The loop structure follows the reference deliberately (it is also the timed CPU baseline): per round,
 * for every product and every i in 0..=d: fold every factor at F::from(i) into FRESH tables, multiply
 * element-wise into a fresh vector, sum; interpolate; sparse-add; serialise; hash; draw the challenge;
 * fold every table at the challenge into fresh tables.  `zkref_set_threads(n)` lets OpenMP split the
 * table-sized loops over n threads (the reference itself is single-threaded; n = 1 is the faithful mode).
 *
 * Field elements cross this file's boundary as canonical little-endian 4 x u64.
 * Each function cites the reference file:line it restates (paths relative to /root/reference).
 */
#include <openssl/sha.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fr; /* Montgomery form, R = 2^256 */

static const uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t NINV = 0xfffffffeffffffffULL; /* -r^-1 mod 2^64 */
static const fr R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};
static int g_threads = 1;

void zkref_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int zkref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int ge_mod(const uint64_t* a) {
    for (int i = 3; i >= 0; i--) { if (a[i] != MOD[i]) return a[i] > MOD[i]; }
    return 1;
}
static void sub_mod(uint64_t* a) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - MOD[i] - br; a[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
}
static fr fr_add(fr a, fr b) {
    fr r; uint64_t c = 0;
    for (int i = 0; i < 4; i++) { u128 s = (u128)a.v[i] + b.v[i] + c; r.v[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    if (c || ge_mod(r.v)) sub_mod(r.v);
    return r;
}
static fr fr_sub(fr a, fr b) {
    fr r; uint64_t br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a.v[i] - b.v[i] - br; r.v[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    if (br) { uint64_t c = 0; for (int i = 0; i < 4; i++) { u128 s = (u128)r.v[i] + MOD[i] + c; r.v[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } }
    return r;
}
/* Montgomery product: full 512-bit schoolbook product, then a separate reduction pass (SOS). */
static fr fr_mul(fr a, fr b) {
    uint64_t t[9] = {0};
    for (int i = 0; i < 4; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) { u128 s = (u128)a.v[i] * b.v[j] + t[i + j] + c; t[i + j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        t[i + 4] = c;
    }
    for (int i = 0; i < 4; i++) {
        uint64_t m = t[i] * NINV, c = 0;
        for (int j = 0; j < 4; j++) { u128 s = (u128)m * MOD[j] + t[i + j] + c; t[i + j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        for (int k = i + 4; k < 9 && c; k++) { u128 s = (u128)t[k] + c; t[k] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    }
    fr r = {{t[4], t[5], t[6], t[7]}};
    if (t[8] || ge_mod(r.v)) sub_mod(r.v);
    return r;
}
static fr fr_from_canon(const uint64_t* c) { fr x = {{c[0], c[1], c[2], c[3]}}; while (ge_mod(x.v)) sub_mod(x.v); return fr_mul(x, R2); }
static void fr_to_canon(fr a, uint64_t* out) { fr one = {{1, 0, 0, 0}}; fr c = fr_mul(a, one); memcpy(out, c.v, 32); }
static fr fr_from_u64(uint64_t x) { uint64_t c[4] = {x, 0, 0, 0}; return fr_from_canon(c); }
static fr fr_zero(void) { fr z = {{0, 0, 0, 0}}; return z; }
static fr fr_one(void) { return fr_from_u64(1); }
static int fr_eq(fr a, fr b) { return memcmp(a.v, b.v, 32) == 0; }
static int fr_is_zero(fr a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
static fr fr_inv(fr a) { /* a^(r-2) */
    uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
    fr acc = fr_one();
    for (int i = 255; i >= 0; i--) { acc = fr_mul(acc, acc); if ((e[i / 64] >> (i % 64)) & 1) acc = fr_mul(acc, a); }
    return acc;
}
static int fr_cmp(fr a, fr b) { /* ark-ff Ord: compares canonical integers */
    uint64_t x[4], y[4]; fr_to_canon(a, x); fr_to_canon(b, y);
    for (int i = 3; i >= 0; i--) { if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1; }
    return 0;
}
/* element.into_bigint().to_bytes_be() -- sumcheck/src/utils.rs:7-9 */
static void fr_be32(fr a, uint8_t* out) {
    uint64_t c[4]; fr_to_canon(a, c);
    for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) out[31 - (8 * i + b)] = (uint8_t)(c[i] >> (8 * b));
}
/* F::from_be_bytes_mod_order -- transcripts/fiat-shamir/src/fiat_shamir.rs:28 */
static fr fr_from_be32(const uint8_t* in) {
    uint64_t c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 32; i++) c[(31 - i) / 8] |= (uint64_t)in[i] << (8 * ((31 - i) % 8));
    return fr_from_canon(c);
}

/* ---- transcripts/fiat-shamir/src/fiat_shamir.rs:5-40 ------------------------------------------- */
typedef struct { SHA256_CTX h; } transcript;
static void tr_new(transcript* t) { SHA256_Init(&t->h); }
static void tr_commit(transcript* t, const uint8_t* d, size_t n) { SHA256_Update(&t->h, d, n); }                 /* :17-19 */
static void tr_challenge(transcript* t, uint8_t out[32]) { SHA256_Final(out, &t->h); SHA256_Init(&t->h); SHA256_Update(&t->h, out, 32); } /* :21-25 */
static fr tr_challenge_field(transcript* t) { uint8_t d[32]; tr_challenge(t, d); return fr_from_be32(d); }      /* :27-29 */

/* ---- polynomial/src/multilinear/evaluation_form.rs ------------------------------------------------- */
typedef struct { size_t n_vars, len; fr* ev; } ml;
static ml ml_alloc(size_t len) { ml m; m.len = len; m.n_vars = 0; while (((size_t)1 << m.n_vars) < len) m.n_vars++; m.ev = (fr*)malloc(len * sizeof(fr)); return m; }
static void ml_free(ml* m) { free(m->ev); m->ev = NULL; }
static ml ml_clone(const ml* a) { ml m = ml_alloc(a->len); memcpy(m.ev, a->ev, a->len * sizeof(fr)); return m; }

/* pick_pairs_with_random_index -- polynomial/src/utils.rs:26-53 (materialises the pair list) */
typedef struct { size_t i, j; } pair_t;
static pair_t* pick_pairs(size_t n, size_t var, size_t* count) {
    pair_t* res = (pair_t*)malloc((n / 2) * sizeof(pair_t));
    size_t k = 0, iters = (size_t)1 << var;
    for (size_t it = 0; it < iters; it++) {
        size_t half = (n / iters) / 2, base = k * 2;
        for (size_t y = 0; y < half; y++) { res[k + y].i = y + base; res[k + y].j = half + y + base; }
        k += half;
    }
    *count = k;
    return res;
}
/* Multilinear::partial_evaluation -- evaluation_form.rs:123-141 : r*y2 + (1-r)*y1 */
static ml ml_partial_evaluation(const ml* a, fr r, size_t var) {
    size_t cnt; pair_t* pairs = pick_pairs(a->len, var, &cnt);
    ml out = ml_alloc(cnt);
    out.n_vars = a->n_vars - 1;
    fr one_minus = fr_sub(fr_one(), r);
#pragma omp parallel for num_threads(g_threads) schedule(static) if (cnt > 4096)
    for (size_t k = 0; k < cnt; k++) out.ev[k] = fr_add(fr_mul(r, a->ev[pairs[k].j]), fr_mul(one_minus, a->ev[pairs[k].i]));
    free(pairs);
    return out;
}
/* Multilinear::evaluation -- :162-175 */
static fr ml_evaluation(const ml* a, const fr* pts) {
    ml cur = ml_clone(a);
    for (size_t i = 0; i < a->n_vars; i++) { ml nx = ml_partial_evaluation(&cur, pts[i], 0); ml_free(&cur); cur = nx; }
    fr r = cur.ev[0]; ml_free(&cur); return r;
}
static fr vec_sum(const fr* v, size_t n) {
    fr total = fr_zero();
#pragma omp parallel num_threads(g_threads) if (n > 4096)
    {
        fr loc = fr_zero();
#pragma omp for schedule(static) nowait
        for (size_t i = 0; i < n; i++) loc = fr_add(loc, v[i]);
#pragma omp critical
        total = fr_add(total, loc);
    }
    return total;
}

/* ---- polynomial/src/composed/composed_multilinear.rs -------------------------------------------------- */
typedef struct { size_t d; ml* polys; } composed;
static composed comp_partial_evaluation(const composed* c, fr r, size_t var) { /* :63-75 */
    composed o; o.d = c->d; o.polys = (ml*)malloc(c->d * sizeof(ml));
    for (size_t k = 0; k < c->d; k++) o.polys[k] = ml_partial_evaluation(&c->polys[k], r, var);
    return o;
}
static void comp_free(composed* c) { for (size_t k = 0; k < c->d; k++) ml_free(&c->polys[k]); free(c->polys); c->polys = NULL; }
static fr* comp_element_wise_product(const composed* c) { /* :105-111 */
    size_t n = c->polys[0].len; fr* out = (fr*)malloc(n * sizeof(fr));
#pragma omp parallel for num_threads(g_threads) schedule(static) if (n > 4096)
    for (size_t i = 0; i < n; i++) { fr p = fr_one(); for (size_t k = 0; k < c->d; k++) p = fr_mul(p, c->polys[k].ev[i]); out[i] = p; }
    return out;
}
/* round-evaluation idiom -- multi_composed_sumcheck.rs:81-89 / composed_sumcheck.rs:41-49 */
static void comp_round_evals(const composed* c, fr* out /* d+1 */) {
    for (size_t i = 0; i <= c->d; i++) {
        composed pe = comp_partial_evaluation(c, fr_from_u64(i), 0);
        fr* prod = comp_element_wise_product(&pe);
        out[i] = vec_sum(prod, pe.polys[0].len);
        free(prod); comp_free(&pe);
    }
}

/* ---- polynomial/src/univariate/sparse_univariate.rs + polynomial/src/utils.rs:78-100 -------------- */
typedef struct { fr coeff, pow; } mono;
typedef struct { size_t n; mono m[2 * 16]; } sparse;
static void lagrange_basis(const fr* xs, size_t n, size_t i, fr* l /* n */) { /* utils.rs:78-100 */
    size_t len = 1; l[0] = fr_one();
    for (size_t j = 0; j < n; j++) if (j != i) {
        fr nl[17]; for (size_t k = 0; k <= len; k++) nl[k] = fr_zero();
        for (size_t k = 0; k < len; k++) { nl[k] = fr_sub(nl[k], fr_mul(l[k], xs[j])); nl[k + 1] = fr_add(nl[k + 1], l[k]); }
        len++; memcpy(l, nl, len * sizeof(fr));
    }
    fr denom = fr_one();
    for (size_t j = 0; j < n; j++) if (j != i) denom = fr_mul(denom, fr_sub(xs[i], xs[j]));
    fr inv = fr_inv(denom);
    for (size_t k = 0; k < n; k++) l[k] = fr_mul(l[k], inv);
}
static sparse sparse_interpolation(const fr* xs, const fr* ys, size_t n) { /* sparse_univariate.rs:40-63 */
    fr result[16];
    for (size_t k = 0; k < n; k++) result[k] = fr_zero();
    for (size_t i = 0; i < n; i++) {
        fr l[17]; lagrange_basis(xs, n, i, l);
        for (size_t k = 0; k < n; k++) result[k] = fr_add(result[k], fr_mul(l[k], ys[i]));
    }
    sparse s; s.n = 0;
    for (size_t k = 0; k < n; k++) if (!fr_is_zero(result[k])) { s.m[s.n].coeff = result[k]; s.m[s.n].pow = fr_from_u64(k); s.n++; } /* :52-60 */
    return s;
}
static sparse sparse_add(const sparse* a, const sparse* b) { /* impl Add :159-203 */
    sparse o; o.n = 0; size_t li = 0, ri = 0;
    while (li < a->n || ri < b->n) {
        if (li < a->n && ri < b->n) {
            int c = fr_cmp(a->m[li].pow, b->m[ri].pow);
            if (c == 0) { o.m[o.n].coeff = fr_add(a->m[li].coeff, b->m[ri].coeff); o.m[o.n].pow = a->m[li].pow; o.n++; li++; ri++; }
            else if (c < 0) o.m[o.n++] = a->m[li++];
            else o.m[o.n++] = b->m[ri++];
        } else if (li < a->n) o.m[o.n++] = a->m[li++];
        else o.m[o.n++] = b->m[ri++];
    }
    return o;
}
static size_t sparse_to_bytes(const sparse* s, uint8_t* out) { /* :27-34 */
    for (size_t i = 0; i < s->n; i++) { fr_be32(s->m[i].coeff, out + 64 * i); fr_be32(s->m[i].pow, out + 64 * i + 32); }
    return 64 * s->n;
}
static fr fr_pow_fr(fr base, fr e) { uint64_t c[4]; fr_to_canon(e, c); fr acc = fr_one();
    for (int i = 255; i >= 0; i--) { acc = fr_mul(acc, acc); if ((c[i / 64] >> (i % 64)) & 1) acc = fr_mul(acc, base); } return acc; }
static fr sparse_evaluate(const sparse* s, fr x) { /* :90-106 */
    fr acc = fr_zero();
    for (size_t i = 0; i < s->n; i++) acc = fr_add(acc, fr_mul(s->m[i].coeff, fr_pow_fr(x, s->m[i].pow)));
    return acc;
}

/* ================================================================================================
 * exported entry points (canonical little-endian 4 x u64 at the boundary)
 * ================================================================================================ */
static ml ml_from_canon(const uint64_t* c, size_t len) {
    ml m = ml_alloc(len);
#pragma omp parallel for num_threads(g_threads) schedule(static) if (len > 4096)
    for (size_t i = 0; i < len; i++) m.ev[i] = fr_from_canon(c + 4 * i);
    return m;
}

/* primitives, for the known-answer tests */
void zkref_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) { fr_to_canon(fr_mul(fr_from_canon(a), fr_from_canon(b)), out); }
void zkref_be32(const uint64_t* a, uint8_t* out) { fr_be32(fr_from_canon(a), out); }
void zkref_partial_evaluation(const uint64_t* evals, uint64_t len, const uint64_t* r, uint64_t var, uint64_t* out) {
    ml a = ml_from_canon(evals, len); ml o = ml_partial_evaluation(&a, fr_from_canon(r), var);
    for (size_t i = 0; i < o.len; i++) fr_to_canon(o.ev[i], out + 4 * i);
    ml_free(&a); ml_free(&o);
}
void zkref_evaluation(const uint64_t* evals, uint64_t len, const uint64_t* pts, uint64_t* out) {
    ml a = ml_from_canon(evals, len); fr p[64];
    for (size_t i = 0; i < a.n_vars; i++) p[i] = fr_from_canon(pts + 4 * i);
    fr_to_canon(ml_evaluation(&a, p), out); ml_free(&a);
}
/* evals at x = 0..n-1 -> monomials (coeff,pow) canonical; returns count */
uint32_t zkref_interpolate(const uint64_t* ys, uint32_t n, uint64_t* out) {
    fr xs[16], y[16];
    for (uint32_t i = 0; i < n; i++) { xs[i] = fr_from_u64(i); y[i] = fr_from_canon(ys + 4 * i); }
    sparse s = sparse_interpolation(xs, y, n);
    for (size_t i = 0; i < s.n; i++) { fr_to_canon(s.m[i].coeff, out + 8 * i); fr_to_canon(s.m[i].pow, out + 8 * i + 4); }
    return (uint32_t)s.n;
}
void zkref_transcript_test(const uint8_t* data, size_t n, uint8_t out1[32], uint8_t out2[32]) {
    transcript t; tr_new(&t); tr_commit(&t, data, n); tr_challenge(&t, out1); tr_challenge(&t, out2);
}

/* protocol: 0 Sumcheck::prove (sumcheck.rs:29-61), 1 ComposedSumcheck::prove (composed_sumcheck.rs:32-67),
 * 2 MultiComposedSumcheckProver::prove_partial, 3 ::prove (multi_composed_sumcheck.rs:47-120).
 * tables: P products, product p has deg[p] tables of 2^n_vars canonical elements, concatenated.
 * Outputs: proof_bytes = concatenation of the per-round transcript messages (== ComposedSumcheckProof::to_bytes
 * for protocols 2/3), *proof_len its length; challenges: n_vars canonical elements.
 * Returns 0. */
int zkref_prove(int protocol, uint32_t n_vars, uint32_t P, const uint32_t* deg, const uint64_t* tables, const uint64_t* sum,
                uint8_t* proof_bytes, size_t* proof_len, uint64_t* challenges) {
    const size_t N = (size_t)1 << n_vars;
    composed* cur = (composed*)malloc(P * sizeof(composed));
    const uint64_t* tp = tables;
    transcript tr; tr_new(&tr);
    if (protocol == 3) { /* transcript.commit(&composed_poly_to_bytes(&poly)) :52 ; Multilinear::to_bytes evaluation_form.rs:54-62 */
        size_t D = 0; for (uint32_t p = 0; p < P; p++) D += deg[p];
        uint8_t* buf = (uint8_t*)malloc(D * N * 32);
        for (size_t i = 0; i < D * N; i++) zkref_be32(tables + 4 * i, buf + 32 * i);
        tr_commit(&tr, buf, D * N * 32); free(buf);
    }
    for (uint32_t p = 0; p < P; p++) { /* current_poly = poly.clone() :72 */
        cur[p].d = deg[p]; cur[p].polys = (ml*)malloc(deg[p] * sizeof(ml));
        for (uint32_t k = 0; k < deg[p]; k++) { cur[p].polys[k] = ml_from_canon(tp, N); tp += 4 * N; }
    }
    if (protocol != 1) { uint8_t b[32]; fr_be32(fr_from_canon(sum), b); tr_commit(&tr, b, 32); } /* :70 / sumcheck.rs:34-35 */
    size_t off = 0;
    for (uint32_t round = 0; round < n_vars; round++) {
        uint8_t msg[64 * 32]; size_t mlen = 0;
        if (protocol == 0) { /* split_poly_into_two_and_sum_each_part evaluation_form.rs:68-74 */
            const ml* t = &cur[0].polys[0]; size_t mid = t->len / 2;
            fr h0 = vec_sum(t->ev, mid), h1 = vec_sum(t->ev + mid, mid);
            fr_be32(h0, msg); fr_be32(h1, msg + 32); mlen = 64;
        } else if (protocol == 1) { /* composed_sumcheck.rs:40-51 */
            fr ev[17]; comp_round_evals(&cur[0], ev);
            for (size_t i = 0; i <= cur[0].d; i++) fr_be32(ev[i], msg + 32 * i);
            mlen = 32 * (cur[0].d + 1);
        } else { /* multi_composed_sumcheck.rs:77-97 */
            sparse round_poly; round_poly.n = 0;
            for (uint32_t p = 0; p < P; p++) {
                fr ev[17], xs[17]; comp_round_evals(&cur[p], ev);
                for (size_t i = 0; i <= cur[p].d; i++) xs[i] = fr_from_u64(i); /* convert_round_poly_to_uni_poly_format utils.rs:29-35 */
                sparse rip = sparse_interpolation(xs, ev, cur[p].d + 1);
                round_poly = sparse_add(&round_poly, &rip);
            }
            mlen = sparse_to_bytes(&round_poly, msg);
        }
        tr_commit(&tr, msg, mlen);
        memcpy(proof_bytes + off, msg, mlen); off += mlen;
        fr r = tr_challenge_field(&tr); /* :99 */
        fr_to_canon(r, challenges + 4 * round);
        for (uint32_t p = 0; p < P; p++) { /* :101-107 */
            composed nx = comp_partial_evaluation(&cur[p], r, 0);
            comp_free(&cur[p]); cur[p] = nx;
        }
    }
    for (uint32_t p = 0; p < P; p++) comp_free(&cur[p]);
    free(cur);
    *proof_len = off;
    return 0;
}

/* sum_p sum_x prod_k f_{p,k}[x]: calculate_poly_sum multi_composed_sumcheck.rs:37-45, composed_sumcheck.rs:28-30 */
void zkref_poly_sum(uint32_t n_vars, uint32_t P, const uint32_t* deg, const uint64_t* tables, uint64_t* out) {
    const size_t N = (size_t)1 << n_vars; const uint64_t* tp = tables; fr total = fr_zero();
    for (uint32_t p = 0; p < P; p++) {
        composed c; c.d = deg[p]; c.polys = (ml*)malloc(deg[p] * sizeof(ml));
        for (uint32_t k = 0; k < deg[p]; k++) { c.polys[k] = ml_from_canon(tp, N); tp += 4 * N; }
        fr* prod = comp_element_wise_product(&c);
        total = fr_add(total, vec_sum(prod, N));
        free(prod); comp_free(&c);
    }
    fr_to_canon(total, out);
}

/* MultiComposedSumcheckVerifier::verify_internal multi_composed_sumcheck.rs:151-181 on serialized monomials.
 * mono: per round `len[r]` (coeff,pow) canonical pairs, concatenated.  Returns 0 ok, -7 verification failed. */
int zkref_verify_partial(uint32_t n_vars, const uint64_t* sum, const uint64_t* mono, const uint32_t* len, uint64_t* sub_sum, uint64_t* challenges) {
    transcript tr; tr_new(&tr);
    fr claimed = fr_from_canon(sum);
    { uint8_t b[32]; fr_be32(claimed, b); tr_commit(&tr, b, 32); }
    const uint64_t* mp = mono;
    for (uint32_t r = 0; r < n_vars; r++) {
        sparse s; s.n = len[r];
        for (uint32_t i = 0; i < len[r]; i++) { s.m[i].coeff = fr_from_canon(mp); s.m[i].pow = fr_from_canon(mp + 4); mp += 8; }
        uint8_t msg[64 * 32]; size_t mlen = sparse_to_bytes(&s, msg);
        tr_commit(&tr, msg, mlen);
        fr c = tr_challenge_field(&tr);
        fr_to_canon(c, challenges + 4 * r);
        fr e01 = fr_add(sparse_evaluate(&s, fr_zero()), sparse_evaluate(&s, fr_one()));
        if (!fr_eq(claimed, e01)) return -7;
        claimed = sparse_evaluate(&s, c);
    }
    fr_to_canon(claimed, sub_sum);
    return 0;
}

/* ---- seeded synthetic tables (same convention as oracle/pymodel.py synth_entry) -------------------- */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31);
}
void zkref_synth_table(uint64_t seed, uint64_t table, uint32_t n_vars, uint64_t* out /* canonical */) {
    const uint64_t base = splitmix64(seed ^ splitmix64(table * 0xD1342543DE82EF95ULL + 0x632BE59BD9B4E019ULL));
    const size_t N = (size_t)1 << n_vars;
#pragma omp parallel for num_threads(g_threads) schedule(static) if (N > 4096)
    for (size_t i = 0; i < N; i++) {
        uint64_t c[4];
        for (int l = 0; l < 4; l++) c[l] = splitmix64(base + ((uint64_t)i * 4 + l) * 0x9E3779B97F4A7C15ULL);
        while (ge_mod(c)) sub_mod(c);
        memcpy(out + 4 * i, c, 32);
    }
}

/* ---- streamed evaluation of a seeded synthetic table (full-size checks: 2^28-entry tables never sit in host memory) ----
 * Multilinear::evaluation (evaluation_form.rs:162-175) of the table zkref_synth_table(seed, table, n_vars) at pts, computed
 * as  sum_i eq(pts, i) * T[i]  with eq(pts, i) = prod_v (bit_v(i) ? pts[v] : 1 - pts[v]), variable 0 = most significant index
 * bit -- the closed form of the reference's n successive folds (tests/test_oracle_cross.py checks the two against each other).
 * The index is split into high and low bits so that only two small eq tables are held. */
static void eq_table(const fr* pts, size_t nv, fr* out /* 2^nv */) {
    out[0] = fr_one();
    size_t len = 1;
    for (size_t v = 0; v < nv; v++) {          /* appending variable v as the new least significant bit keeps MSB-first order */
        fr one_minus = fr_sub(fr_one(), pts[v]);
        for (size_t i = len; i-- > 0;) { fr e = out[i]; out[2 * i] = fr_mul(e, one_minus); out[2 * i + 1] = fr_mul(e, pts[v]); }
        len *= 2;
    }
}
void zkref_eval_synth(uint64_t seed, uint64_t table, uint32_t n_vars, const uint64_t* pts_canon, uint64_t* out) {
    const uint64_t base = splitmix64(seed ^ splitmix64(table * 0xD1342543DE82EF95ULL + 0x632BE59BD9B4E019ULL));
    fr pts[64];
    for (uint32_t i = 0; i < n_vars; i++) pts[i] = fr_from_canon(pts_canon + 4 * i);
    const uint32_t nl = n_vars < 12 ? n_vars : 12, nh = n_vars - nl;
    const size_t NH = (size_t)1 << nh, NLO = (size_t)1 << nl;
    fr* eq_hi = (fr*)malloc(NH * sizeof(fr));
    fr* eq_lo = (fr*)malloc(NLO * sizeof(fr));
    eq_table(pts, nh, eq_hi);
    eq_table(pts + nh, nl, eq_lo);
    fr total = fr_zero();
#pragma omp parallel num_threads(g_threads) if (NH > 1)
    {
        fr loc = fr_zero();
#pragma omp for schedule(static) nowait
        for (size_t hi = 0; hi < NH; hi++) {
            fr inner = fr_zero();
            for (size_t lo = 0; lo < NLO; lo++) {
                const uint64_t i = (uint64_t)hi * NLO + lo;
                uint64_t c[4];
                for (int l = 0; l < 4; l++) c[l] = splitmix64(base + (i * 4 + l) * 0x9E3779B97F4A7C15ULL);
                while (ge_mod(c)) sub_mod(c);
                inner = fr_add(inner, fr_mul(eq_lo[lo], fr_from_canon(c)));
            }
            loc = fr_add(loc, fr_mul(eq_hi[hi], inner));
        }
#pragma omp critical
        total = fr_add(total, loc);
    }
    free(eq_hi); free(eq_lo);
    fr_to_canon(total, out);
}
