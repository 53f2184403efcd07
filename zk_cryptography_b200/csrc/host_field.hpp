// Host-side BLS12-381 Fr, SHA-256, Fiat-Shamir transcript and sparse univariate polynomial.
//
// north_star keeps the Fiat-Shamir transcript -- and with it the per-round interpolation, sparse
// addition, serialisation and challenge derivation (a handful of field elements per round) -- on the
// host.  This header is that host logic; the table-sized work (round evaluations, folds) is CUDA only.
//
// Mirrors, by name and behaviour:
//   FiatShamirTranscript           transcripts/fiat-shamir/src/fiat_shamir.rs:5-40
//   SparseUnivariatePolynomial     polynomial/src/univariate/sparse_univariate.rs:11-203
//   lagrange interpolation         polynomial/src/utils.rs:78-100 (same values, Newton-free direct form)
//   convert_field_to_byte etc.     sumcheck/src/utils.rs:7-59
// Field semantics are ark-ff 0.4.2 Fp256<MontBackend<FrConfig,4>> (third-party, not in the tree):
// 4 x u64 little-endian limbs, Montgomery form, R = 2^256.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define ZKSC_HAVE_SHANI 1
#else
#define ZKSC_HAVE_SHANI 0
#endif

namespace zksc {
namespace host {

typedef unsigned __int128 u128;

struct FrH {
    uint64_t v[4];  // Montgomery form
    bool operator==(const FrH& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    bool operator!=(const FrH& o) const { return !(*this == o); }
};

static const uint64_t kMod[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const uint64_t kInv = 0xfffffffeffffffffull;  // -r^-1 mod 2^64
static const FrH kR2 = {{0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull}};
static const FrH kOne = {{0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full}};
static const FrH kZero = {{0, 0, 0, 0}};

inline bool geq_mod(const uint64_t* a) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > kMod[i]) return true;
        if (a[i] < kMod[i]) return false;
    }
    return true;
}
inline void sub_mod_inplace(uint64_t* a) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - kMod[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}
inline FrH add(const FrH& a, const FrH& b) {
    FrH r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (c || geq_mod(r.v)) sub_mod_inplace(r.v);
    return r;
}
inline FrH sub(const FrH& a, const FrH& b) {
    FrH r;
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.v[i] - b.v[i] - borrow;
        r.v[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)r.v[i] + kMod[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
inline FrH neg(const FrH& a) { return sub(kZero, a); }
// Montgomery product (CIOS on 64-bit limbs)
inline FrH mul(const FrH& a, const FrH& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.v[j] * b.v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * kInv;
        c = ((u128)m * kMod[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * kMod[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    FrH r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(r.v)) sub_mod_inplace(r.v);
    return r;
}
inline FrH from_canonical(const uint64_t c[4]) {  // c < 2^256, any representative
    FrH x = {{c[0], c[1], c[2], c[3]}};
    while (geq_mod(x.v)) sub_mod_inplace(x.v);
    return mul(x, kR2);
}
inline void to_canonical(const FrH& a, uint64_t out[4]) {
    FrH one = {{1, 0, 0, 0}};
    FrH c = mul(a, one);
    memcpy(out, c.v, 32);
}
inline FrH from_u64(uint64_t x) {
    uint64_t c[4] = {x, 0, 0, 0};
    return from_canonical(c);
}
// Shift table of a challenge for the device's fixed-multiplicand fold (fr.cuh mul_fixed_rows):
// w[i] = r * 2^(32 (i + 2)) mod p as 8 little-endian 32-bit limbs, r = the plain value of r_mont.
inline void fold_table(const FrH& r_mont, uint32_t w[8][8]) {
    // 2^(32 e) mod p as RAW limbs, e = 2..9 (constants, built once): mul(r R, X) = r X
    static const std::vector<FrH> X = [] {
        std::vector<FrH> x(8, kZero);
        for (int i = 0; i < 8; i++) {
            const int e = i + 2;
            if (e < 8) x[i].v[e / 2] = 1ull << (32 * (e % 2));
            else if (e == 8) x[i] = kOne;              // 2^256 mod p
            else x[i] = from_u64(1ull << 32);          // 2^288 mod p
        }
        return x;
    }();
    for (int i = 0; i < 8; i++) {
        FrH v = mul(r_mont, X[i]);
        memcpy(w[i], v.v, 32);
    }
}
inline FrH pow_u(const FrH& a, const uint64_t e[4]) {
    FrH acc = kOne;
    for (int i = 255; i >= 0; i--) {
        acc = mul(acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, a);
    }
    return acc;
}
inline FrH inverse(const FrH& a) {  // a != 0 ; a^(r-2)
    uint64_t e[4] = {kMod[0] - 2, kMod[1], kMod[2], kMod[3]};
    return pow_u(a, e);
}
// element.into_bigint().to_bytes_be()  (sumcheck/src/utils.rs:7-9)
inline void to_be_bytes(const FrH& a, uint8_t out[32]) {
    uint64_t c[4];
    to_canonical(a, c);
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 8; b++) out[31 - (i * 8 + b)] = (uint8_t)(c[i] >> (8 * b));
}
// F::from_be_bytes_mod_order on exactly 32 bytes (transcripts/fiat-shamir/src/fiat_shamir.rs:28)
inline FrH from_be_bytes_mod_order(const uint8_t in[32]) {
    uint64_t c[4] = {0, 0, 0, 0};
    for (int i = 0; i < 32; i++) c[(31 - i) / 8] |= (uint64_t)in[i] << (8 * ((31 - i) % 8));
    return from_canonical(c);
}
// canonical comparison (ark-ff Ord on Fp compares into_bigint())
inline int cmp_canonical(const FrH& a, const FrH& b) {
    uint64_t x[4], y[4];
    to_canonical(a, x);
    to_canonical(b, y);
    for (int i = 3; i >= 0; i--) {
        if (x[i] < y[i]) return -1;
        if (x[i] > y[i]) return 1;
    }
    return 0;
}

// What the round kernels deliver per product (kernels.cuh accumulate_points) -> the evaluations h(0..d) the protocol speaks of.
// `e` holds degree + 1 elements (4 x u64 each), slots 0 and 1 = h(0), h(1) for every degree.
//   degree 2: slot 2 = the leading coefficient h(inf);            h(2) = 2 h(1) - h(0) + 2 h(inf)
//   degree 3: slots 2, 3 = h(-1), h(inf).  With h(t) = c0 + c1 t + c2 t^2 + c3 t^3:
//             c0 = h(0), c3 = h(inf), c2 = (h(1) + h(-1)) / 2 - c0, c1 = (h(1) - h(-1)) / 2 - c3,
//             h(2) = c0 + 2 c1 + 4 c2 + 8 c3,  h(3) = c0 + 3 c1 + 9 c2 + 27 c3
//   other degrees: the slots are the evaluations already.
// Exact field arithmetic: the canonical values are those of evaluating at 2 (and 3) directly.
inline void round_slots_to_evals(uint32_t degree, uint64_t* e) {
    auto ld = [](const uint64_t* p) { FrH h; memcpy(h.v, p, 32); return h; };
    auto st = [](uint64_t* p, const FrH& h) { memcpy(p, h.v, 32); };
    if (degree == 3) {
        static const FrH inv2 = inverse(from_u64(2));
        static const FrH k3 = from_u64(3), k4 = from_u64(4), k8 = from_u64(8), k9 = from_u64(9), k27 = from_u64(27);
        const FrH c0 = ld(e), h1 = ld(e + 4), hm = ld(e + 8), c3 = ld(e + 12);
        const FrH c2 = sub(mul(add(h1, hm), inv2), c0);
        const FrH c1 = sub(mul(sub(h1, hm), inv2), c3);
        const FrH c1_2 = add(c1, c1);
        st(e + 8, add(add(c0, c1_2), add(mul(c2, k4), mul(c3, k8))));
        st(e + 12, add(add(c0, mul(c1, k3)), add(mul(c2, k9), mul(c3, k27))));
    } else if (degree == 2) {
        const FrH h0 = ld(e), h1 = ld(e + 4), hi = ld(e + 8);
        const FrH s = add(sub(h1, h0), hi);       // h(1) - h(0) + h(inf)
        st(e + 8, add(add(s, s), h0));            // 2 (h(1) - h(0) + h(inf)) + h(0)
    }
}

// ------------------------------------------------------------------------------------------------
// SHA-256 (FIPS 180-4), streaming.  Replaces sha2 0.10.8 `Sha256::{new,update,finalize_reset}`.
// ------------------------------------------------------------------------------------------------
class Sha256 {
   public:
    Sha256() { reset(); }
    void reset() {
        static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        memcpy(h_, iv, sizeof iv);
        len_ = 0;
        fill_ = 0;
    }
    void update(const uint8_t* p, size_t n) {
        len_ += n;
        if (fill_) {
            size_t take = 64 - fill_ < n ? 64 - fill_ : n;
            memcpy(buf_ + fill_, p, take);
            fill_ += take; p += take; n -= take;
            if (fill_ == 64) { block(buf_); fill_ = 0; }
        }
        while (n >= 64) { block(p); p += 64; n -= 64; }
        if (n) { memcpy(buf_, p, n); fill_ = n; }
    }
    void finalize_reset(uint8_t out[32]) {
        uint64_t bits = len_ * 8;
        uint8_t pad[72];
        size_t padlen = (fill_ < 56) ? 56 - fill_ : 120 - fill_;
        memset(pad, 0, sizeof pad);
        pad[0] = 0x80;
        for (int i = 0; i < 8; i++) pad[padlen + i] = (uint8_t)(bits >> (56 - 8 * i));
        update(pad, padlen + 8);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(h_[i] >> 24); out[4 * i + 1] = (uint8_t)(h_[i] >> 16);
            out[4 * i + 2] = (uint8_t)(h_[i] >> 8); out[4 * i + 3] = (uint8_t)h_[i];
        }
        reset();
    }

   private:
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    static const uint32_t* round_constants() {
        alignas(16) static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
            0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
            0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
            0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
            0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
            0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        return K;
    }
#if ZKSC_HAVE_SHANI
    // One compression with the x86 SHA extensions (sha256rnds2 / sha256msg1 / sha256msg2): the per-round host step of every proof
    // hashes 3-5 blocks, and with 64 batched proofs per launch (c5) the scalar compression was a visible share of the step.
    // Message schedule per group g of four rounds: W[g] = msg2(msg1(W[g-4], W[g-3]) + alignr(W[g-1], W[g-2], 4), W[g-1]).
    __attribute__((target("sha,sse4.1,ssse3"))) static void block_shani(uint32_t h[8], const uint8_t* p) {
        const uint32_t* K = round_constants();
        const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
        __m128i tmp = _mm_loadu_si128((const __m128i*)&h[0]);       // DCBA
        __m128i st1 = _mm_loadu_si128((const __m128i*)&h[4]);       // HGFE
        tmp = _mm_shuffle_epi32(tmp, 0xB1);                         // CDAB
        st1 = _mm_shuffle_epi32(st1, 0x1B);                         // EFGH
        __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                 // ABEF
        st1 = _mm_blend_epi16(st1, tmp, 0xF0);                      // CDGH
        const __m128i save0 = st0, save1 = st1;
        __m128i w[16];
        for (int g = 0; g < 16; g++) {
            if (g < 4) {
                w[g] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * g)), bswap);
            } else {
                __m128i x = _mm_sha256msg1_epu32(w[g - 4], w[g - 3]);
                x = _mm_add_epi32(x, _mm_alignr_epi8(w[g - 1], w[g - 2], 4));
                w[g] = _mm_sha256msg2_epu32(x, w[g - 1]);
            }
            __m128i m = _mm_add_epi32(w[g], _mm_load_si128((const __m128i*)&K[4 * g]));
            st1 = _mm_sha256rnds2_epu32(st1, st0, m);
            m = _mm_shuffle_epi32(m, 0x0E);
            st0 = _mm_sha256rnds2_epu32(st0, st1, m);
        }
        st0 = _mm_add_epi32(st0, save0);
        st1 = _mm_add_epi32(st1, save1);
        tmp = _mm_shuffle_epi32(st0, 0x1B);                         // FEBA
        st1 = _mm_shuffle_epi32(st1, 0xB1);                         // DCHG
        st0 = _mm_blend_epi16(tmp, st1, 0xF0);                      // DCBA
        st1 = _mm_alignr_epi8(st1, tmp, 8);                         // HGFE
        _mm_storeu_si128((__m128i*)&h[0], st0);
        _mm_storeu_si128((__m128i*)&h[4], st1);
    }
    static bool have_shani() {
        static const bool ok = __builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3") && !getenv("ZKSC_NO_SHANI");
        return ok;
    }
#endif
    void block(const uint8_t* p) {
#if ZKSC_HAVE_SHANI
        if (have_shani()) { block_shani(h_, p); return; }
#endif
        const uint32_t* K = round_constants();
        
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h_[0], b = h_[1], c = h_[2], d = h_[3], e = h_[4], f = h_[5], g = h_[6], h = h_[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = h + S1 + ch + K[i] + w[i];
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h_[0] += a; h_[1] += b; h_[2] += c; h_[3] += d; h_[4] += e; h_[5] += f; h_[6] += g; h_[7] += h;
    }
    uint32_t h_[8];
    uint64_t len_;
    uint8_t buf_[64];
    size_t fill_;
};

// transcripts/fiat-shamir/src/fiat_shamir.rs:5-40
class FiatShamirTranscript {
   public:
    void commit(const uint8_t* data, size_t n) { hasher_.update(data, n); }                       // :17-19
    void commit(const std::vector<uint8_t>& d) { hasher_.update(d.data(), d.size()); }
    void commit_field(const FrH& x) { uint8_t b[32]; to_be_bytes(x, b); hasher_.update(b, 32); }
    void challenge(uint8_t out[32]) { hasher_.finalize_reset(out); hasher_.update(out, 32); }     // :21-25
    FrH evaluate_challenge_into_field() { uint8_t d[32]; challenge(d); return from_be_bytes_mod_order(d); }  // :27-29
    std::vector<FrH> evaluate_n_challenge_into_field(size_t n) {                                  // :31-39
        std::vector<FrH> r;
        for (size_t i = 0; i < n; i++) r.push_back(evaluate_challenge_into_field());
        return r;
    }

   private:
    Sha256 hasher_;
};

// polynomial/src/univariate/sparse_univariate.rs
struct UnivariateMonomial {
    FrH coeff, pow;
};
struct SparseUnivariatePolynomial {
    std::vector<UnivariateMonomial> monomial;

    static SparseUnivariatePolynomial zero() { return SparseUnivariatePolynomial(); }  // :23-25

    // :27-34
    void to_bytes(std::vector<uint8_t>& out) const {
        for (const auto& m : monomial) {
            uint8_t b[64];
            to_be_bytes(m.coeff, b);
            to_be_bytes(m.pow, b + 32);
            out.insert(out.end(), b, b + 64);
        }
    }

    // :40-63 with lagrange_basis polynomial/src/utils.rs:78-100.  Same canonical coefficients (the
    // interpolant is unique); monomials whose coefficient is zero are dropped (:52-60).
    static SparseUnivariatePolynomial interpolation(const std::vector<FrH>& xs, const std::vector<FrH>& ys) {
        const size_t n = xs.size();
        std::vector<FrH> result(n, kZero);
        for (size_t i = 0; i < n; i++) {
            std::vector<FrH> l(1, kOne);
            FrH denom = kOne;
            for (size_t j = 0; j < n; j++) {
                if (j == i) continue;
                std::vector<FrH> nl(l.size() + 1, kZero);
                for (size_t k = 0; k < l.size(); k++) {
                    nl[k] = sub(nl[k], mul(l[k], xs[j]));
                    nl[k + 1] = add(nl[k + 1], l[k]);
                }
                l.swap(nl);
                denom = mul(denom, sub(xs[i], xs[j]));
            }
            FrH scale = mul(inverse(denom), ys[i]);
            for (size_t k = 0; k < l.size(); k++) result[k] = add(result[k], mul(l[k], scale));
        }
        SparseUnivariatePolynomial p;
        for (size_t k = 0; k < n; k++)
            if (result[k] != kZero) p.monomial.push_back({result[k], from_u64(k)});
        return p;
    }
    // Dense coefficients of the interpolant through (0, ys[0]) .. (d, ys[d]).  Same field values as
    // `interpolation` (the interpolant is unique); the d+1 inverse denominators prod_{j != i}(i - j) depend
    // on d only and are computed once (the general routine pays one Fermat inversion per basis polynomial,
    // ~10 us each -- too slow for a per-round host step).
    static std::vector<FrH> dense_interpolate_evals(const std::vector<FrH>& ys) {
        const size_t n = ys.size();
        // Degrees 1-3 in closed form (Newton's forward differences over x = 0, 1, 2, 3): the per-round host step of every prover here.
        // The interpolant is unique, so these are the field values of the general routine below (tests/test_host_logic.py compares them).
        if (n == 2) return {ys[0], sub(ys[1], ys[0])};
        if (n == 3 || n == 4) {
            static const FrH inv2 = inverse(from_u64(2)), inv6 = inverse(from_u64(6));
            const FrH d1 = sub(ys[1], ys[0]);
            const FrH d2 = add(sub(ys[2], add(ys[1], ys[1])), ys[0]);                 // y2 - 2 y1 + y0
            const FrH h2 = mul(d2, inv2);                                             // coefficient of x (x - 1)
            if (n == 3) return {ys[0], sub(d1, h2), h2};
            // y3 - 3 y2 + 3 y1 - y0
            const FrH t21 = sub(ys[1], ys[2]);
            const FrH d3 = add(sub(ys[3], ys[0]), add(t21, add(t21, t21)));
            const FrH c3 = mul(d3, inv6);                                             // coefficient of x (x - 1) (x - 2) = x^3 - 3 x^2 + 2 x
            const FrH c3_2 = add(c3, c3);
            return {ys[0], add(sub(d1, h2), c3_2), sub(h2, add(c3_2, c3)), c3};
        }
        if (n >= 64) {
            std::vector<FrH> xs;
            for (size_t i = 0; i < n; i++) xs.push_back(from_u64(i));
            SparseUnivariatePolynomial sp = interpolation(xs, ys);
            std::vector<FrH> dense(n, kZero);
            for (const auto& m : sp.monomial) { uint64_t e[4]; to_canonical(m.pow, e); dense[e[0]] = m.coeff; }
            return dense;
        }
        // basis[n][i] = the coefficients of the i-th Lagrange basis polynomial over x = 0..n-1: depends on n only, built once per
        // thread (one Fermat inversion per basis polynomial, ~10 us each -- far too slow for a per-round host step); a round then
        // costs n^2 multiplications.
        static thread_local std::vector<std::vector<std::vector<FrH>>> basis(64);
        std::vector<std::vector<FrH>>& B = basis[n];
        if (B.size() != n) {
            B.assign(n, std::vector<FrH>());
            std::vector<FrH> l, nl;
            for (size_t i = 0; i < n; i++) {
                FrH denom = kOne;
                l.assign(1, kOne);
                for (size_t j = 0; j < n; j++) {
                    if (j == i) continue;
                    const FrH xj = from_u64(j);
                    denom = mul(denom, sub(from_u64(i), xj));
                    nl.assign(l.size() + 1, kZero);
                    for (size_t k = 0; k < l.size(); k++) {
                        nl[k] = sub(nl[k], mul(l[k], xj));
                        nl[k + 1] = add(nl[k + 1], l[k]);
                    }
                    l.swap(nl);
                }
                const FrH inv = inverse(denom);
                B[i].resize(n);
                for (size_t k = 0; k < n; k++) B[i][k] = mul(l[k], inv);
            }
        }
        std::vector<FrH> result(n, kZero);
        for (size_t i = 0; i < n; i++)
            for (size_t k = 0; k < n; k++) result[k] = add(result[k], mul(B[i][k], ys[i]));
        return result;
    }
    // F::from(k) for the small exponents of a round polynomial, Montgomery form, built once
    static const FrH& small_mont(size_t k) {
        static const std::vector<FrH> tab = [] {
            std::vector<FrH> t;
            for (uint64_t i = 0; i < 64; i++) t.push_back(from_u64(i));
            return t;
        }();
        return tab[k];
    }
    // evaluations at x = 0..d  (sumcheck/src/utils.rs:29-35 convert_round_poly_to_uni_poly_format) ->
    // SparseUnivariatePolynomial::interpolation (:40-63): monomials with a zero coefficient are dropped (:52-60)
    static SparseUnivariatePolynomial interpolate_evals(const std::vector<FrH>& ys) {
        std::vector<FrH> dense = dense_interpolate_evals(ys);
        SparseUnivariatePolynomial p;
        for (size_t k = 0; k < dense.size(); k++)
            if (dense[k] != kZero) p.monomial.push_back({dense[k], k < 64 ? small_mont(k) : from_u64(k)});
        return p;
    }
    // value at `point` of the interpolant through (i, ys[i]): Horner on the dense coefficients
    static FrH evaluate_evals_at(const std::vector<FrH>& ys, const FrH& point) {
        std::vector<FrH> c = dense_interpolate_evals(ys);
        FrH acc = kZero;
        for (size_t k = c.size(); k-- > 0;) acc = add(mul(acc, point), c[k]);
        return acc;
    }

    // :90-106  sum coeff * point^pow (pow is a field element used as a 256-bit exponent)
    FrH evaluate(const FrH& point) const {
        FrH acc = kZero;
        for (const auto& m : monomial) {
            uint64_t e[4];
            to_canonical(m.pow, e);
            acc = add(acc, mul(m.coeff, pow_u(point, e)));
        }
        return acc;
    }

    // impl Add :159-203 -- ordered merge by pow; equal powers are summed and KEPT even when zero
    SparseUnivariatePolynomial operator+(const SparseUnivariatePolynomial& rhs) const {
        SparseUnivariatePolynomial out;
        size_t li = 0, ri = 0;
        while (li < monomial.size() || ri < rhs.monomial.size()) {
            if (li < monomial.size() && ri < rhs.monomial.size()) {
                const auto &l = monomial[li], &r = rhs.monomial[ri];
                int c = cmp_canonical(l.pow, r.pow);
                if (c == 0) { out.monomial.push_back({add(l.coeff, r.coeff), l.pow}); li++; ri++; }
                else if (c < 0) { out.monomial.push_back(l); li++; }
                else { out.monomial.push_back(r); ri++; }
            } else if (li < monomial.size()) out.monomial.push_back(monomial[li++]);
            else out.monomial.push_back(rhs.monomial[ri++]);
        }
        return out;
    }
};

}  // namespace host
}  // namespace zksc
