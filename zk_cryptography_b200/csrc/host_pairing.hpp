// BLS12-381 pairing on the host, natively: the pairing check of MultilinearKZG::verify (kzg/src/multilinear_kzg.rs:90-116;
// sum_pairing_results, kzg/src/utils.rs:42-61) and of SuccintGKRProtocol::verify (gkr/src/succint_protocol.rs:169-266).
// The reference takes `Bls12_381::pairing` from ark-ec 0.4.2 (third party, absent from /root/reference).  A verifier computes n + 1
// pairings of single points -- constant-size host work like the transcript; nothing here is table-sized, nothing here runs on the GPU.
//
// The same textbook construction as zk_cryptography_b200/pairing.py (which stays as the independent cross-check of the tests), chosen for
// being easy to check rather than fast:  Fq2 = Fq[u]/(u^2 + 1), Fq6 = Fq2[v]/(v^3 - (1 + u)), Fq12 = Fq6[w]/(w^2 - v);  G2 points are
// mapped into E(Fq12): y^2 = x^3 + 4 by (x, y) -> (x / w^2, y / w^3);  e(P, Q) = conj(f_{|x|, Q}(P)) ^ ((p^12 - 1) / r) with the affine
// Miller loop (vertical lines dropped: they lie in Fq6 and die in the final exponentiation) and a plain square-and-multiply final
// exponentiation.  ~5 ms per Miller loop, ~20 ms per final exponentiation.
#pragma once
#include <cstdint>
#include <cstring>

namespace zksc {
namespace pairing {

typedef unsigned __int128 u128;
struct Fq { uint64_t v[6]; };
static const uint64_t kP[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const uint64_t kPInv = 0x89f3fffcfffcfffdull;      // -p^-1 mod 2^64
static const Fq kR2 = {{0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull}};
static const Fq kOne = {{0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull}};
static const Fq kZero = {{0, 0, 0, 0, 0, 0}};
static const uint64_t kPm2[6] = {0xb9feffffffffaaa9ull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const uint64_t kXAbs = 0xd201000000010000ull;      // |x| of the curve parameter x = -|x|
// (p^12 - 1) / r, little-endian 64-bit words (4314 bits)
static const uint64_t kFinalExp[68] = {
    0xc0bcb9b55df57510ull, 0x25f98630e68bfb24ull, 0x4406fbc8fbd5f489ull, 0x8e2f8491d12191a0ull,
    0x3e9d71650a6f8069ull, 0x226c2f011d4cab80ull, 0x67f67c4717489119ull, 0xaf3f881bd88592d7ull,
    0x1a67e49eeed2161dull, 0xe5b78c7869aeb218ull, 0xf6539314043f7bbcull, 0x73f62537f2701aaeull,
    0xaff1c910e9622d2aull, 0x6283313492caa9d4ull, 0x2e2f3ec2bea83d19ull, 0xa4c7e79fb02faa73ull,
    0x6c49637fd7961be1ull, 0x08e88adce8817745ull, 0x35de3f7a36399917ull, 0x9c1d9f7c31759c36ull,
    0xfa9e13c24ea820b0ull, 0x3fc56947a403577dull, 0xa4c1b6dcfc5cceb7ull, 0x1bbd81367066bca6ull,
    0x0418a3ef0bc62775ull, 0x49bf9b71a9f9e010ull, 0x511291097db60b17ull, 0x498345c6e5308f1cull,
    0x6d8823b19dadd7c2ull, 0x92004cedd556952cull, 0x4c6bec3ec03ef195ull, 0x0a1fad20044ce6adull,
    0xc55d3109cd15948dull, 0x334f46c02c3f0bd0ull, 0x3b5a62eb34c05739ull, 0x724538411d1676a5ull,
    0x127a1b5ad0463434ull, 0x61a474c5c85b0129ull, 0x8dfc8e2886ef965eull, 0x96532fef459f1243ull,
    0x40ee7169cdc10412ull, 0x9c40a68eb74bb22aull, 0x25118790f4684d0bull, 0x596bc293c8d4c01full,
    0x1064837f27611212ull, 0x077ffb10bf24dde4ull, 0xc49f570bcd2b01f3ull, 0x1a0c5bf24c374693ull,
    0x350da5359bc73ab6ull, 0xd2670d93e4d7acddull, 0xd39099b86e1ab656ull, 0x19328148978e2b0dull,
    0xb113f414386b0e88ull, 0x07a0dce2630d9aa4ull, 0xa927e7bb93753318ull, 0xe347aa68ad49466full,
    0x1c0ad0d6106feaf4ull, 0xc872ee83ff3a0f0full, 0x074e43b9a660835cull, 0xc0aadff5e9cfee9aull,
    0x30698e8cc7deada9ull, 0xd1073776ab353f2cull, 0x17848517badc3a43ull, 0x7363baa13f8d14a9ull,
    0xd4977b3f7d4507d0ull, 0x496a1c0a89ee0193ull, 0xdcc825b7e1bda9c0ull, 0x0000000002ee1db5ull};

inline bool geq_p(const uint64_t a[6]) {
    for (int i = 5; i >= 0; i--) {
        if (a[i] > kP[i]) return true;
        if (a[i] < kP[i]) return false;
    }
    return true;
}
inline void sub_p(uint64_t a[6]) {
    u128 b = 0;
    for (int i = 0; i < 6; i++) {
        const u128 d = (u128)a[i] - kP[i] - (uint64_t)b;
        a[i] = (uint64_t)d;
        b = (d >> 64) & 1;
    }
}
inline bool is_zero(const Fq& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5]) == 0; }
inline bool eq(const Fq& a, const Fq& b) { return memcmp(a.v, b.v, 48) == 0; }
inline Fq add(const Fq& a, const Fq& b) {
    Fq r;
    u128 c = 0;
    for (int i = 0; i < 6; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    if (c || geq_p(r.v)) sub_p(r.v);          // p < 2^381: no carry out of 384 bits in fact
    return r;
}
inline Fq sub(const Fq& a, const Fq& b) {
    Fq r;
    u128 bw = 0;
    for (int i = 0; i < 6; i++) {
        const u128 d = (u128)a.v[i] - b.v[i] - (uint64_t)bw;
        r.v[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    if (bw) {
        u128 c = 0;
        for (int i = 0; i < 6; i++) { c += (u128)r.v[i] + kP[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
}
inline Fq neg(const Fq& a) { return is_zero(a) ? a : sub(kZero, a); }
inline Fq mul(const Fq& a, const Fq& b) {     // Montgomery product, CIOS
    uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
        u128 c = 0;
        for (int j = 0; j < 6; j++) { c += (u128)a.v[j] * b.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[6];
        t[6] = (uint64_t)c;
        t[7] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * kPInv;
        c = ((u128)m * kP[0] + t[0]) >> 64;
        for (int j = 1; j < 6; j++) { c += (u128)m * kP[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[6];
        t[5] = (uint64_t)c;
        t[6] = t[7] + (uint64_t)(c >> 64);
    }
    Fq r;
    memcpy(r.v, t, 48);
    if (t[6] || geq_p(r.v)) sub_p(r.v);
    return r;
}
inline Fq from_canonical(const uint64_t c[6]) { Fq x; memcpy(x.v, c, 48); return mul(x, kR2); }
inline void to_canonical(const Fq& a, uint64_t out[6]) { Fq one = {{1, 0, 0, 0, 0, 0}}; const Fq c = mul(a, one); memcpy(out, c.v, 48); }
inline Fq from_u64(uint64_t x) { const uint64_t c[6] = {x, 0, 0, 0, 0, 0}; return from_canonical(c); }
inline Fq inv(const Fq& a) {                  // a^(p-2), a != 0
    Fq acc = kOne;
    for (int i = 383; i >= 0; i--) {
        acc = mul(acc, acc);
        if ((kPm2[i / 64] >> (i % 64)) & 1) acc = mul(acc, a);
    }
    return acc;
}

struct Fq2 { Fq c0, c1; };                    // c0 + c1 u, u^2 = -1
inline Fq2 add(const Fq2& a, const Fq2& b) { return {add(a.c0, b.c0), add(a.c1, b.c1)}; }
inline Fq2 sub(const Fq2& a, const Fq2& b) { return {sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
inline Fq2 neg(const Fq2& a) { return {neg(a.c0), neg(a.c1)}; }
inline bool is_zero(const Fq2& a) { return is_zero(a.c0) && is_zero(a.c1); }
inline bool eq(const Fq2& a, const Fq2& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
inline Fq2 mul(const Fq2& a, const Fq2& b) {
    const Fq t0 = mul(a.c0, b.c0), t1 = mul(a.c1, b.c1);
    return {sub(t0, t1), sub(sub(mul(add(a.c0, a.c1), add(b.c0, b.c1)), t0), t1)};
}
inline Fq2 mul_xi(const Fq2& a) { return {sub(a.c0, a.c1), add(a.c0, a.c1)}; }      // times the non-residue 1 + u
inline Fq2 inv(const Fq2& a) {
    const Fq d = inv(add(mul(a.c0, a.c0), mul(a.c1, a.c1)));
    return {mul(a.c0, d), neg(mul(a.c1, d))};
}
static const Fq2 kZero2 = {kZero, kZero};
static const Fq2 kOne2 = {kOne, kZero};

struct Fq6 { Fq2 c0, c1, c2; };               // c0 + c1 v + c2 v^2, v^3 = 1 + u
inline Fq6 add(const Fq6& a, const Fq6& b) { return {add(a.c0, b.c0), add(a.c1, b.c1), add(a.c2, b.c2)}; }
inline Fq6 sub(const Fq6& a, const Fq6& b) { return {sub(a.c0, b.c0), sub(a.c1, b.c1), sub(a.c2, b.c2)}; }
inline Fq6 neg(const Fq6& a) { return {neg(a.c0), neg(a.c1), neg(a.c2)}; }
inline bool eq(const Fq6& a, const Fq6& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1) && eq(a.c2, b.c2); }
inline Fq6 mul(const Fq6& a, const Fq6& b) {
    const Fq2 t00 = mul(a.c0, b.c0), t11 = mul(a.c1, b.c1), t22 = mul(a.c2, b.c2);
    return {add(t00, mul_xi(add(mul(a.c1, b.c2), mul(a.c2, b.c1)))), add(add(mul(a.c0, b.c1), mul(a.c1, b.c0)), mul_xi(t22)),
            add(add(mul(a.c0, b.c2), mul(a.c2, b.c0)), t11)};
}
inline Fq6 mul_by_v(const Fq6& a) { return {mul_xi(a.c2), a.c0, a.c1}; }
inline Fq6 inv(const Fq6& a) {
    const Fq2 t0 = sub(mul(a.c0, a.c0), mul_xi(mul(a.c1, a.c2)));
    const Fq2 t1 = sub(mul_xi(mul(a.c2, a.c2)), mul(a.c0, a.c1));
    const Fq2 t2 = sub(mul(a.c1, a.c1), mul(a.c0, a.c2));
    const Fq2 d = inv(add(mul(a.c0, t0), mul_xi(add(mul(a.c2, t1), mul(a.c1, t2)))));
    return {mul(t0, d), mul(t1, d), mul(t2, d)};
}
static const Fq6 kZero6 = {kZero2, kZero2, kZero2};
static const Fq6 kOne6 = {kOne2, kZero2, kZero2};

struct Fq12 { Fq6 c0, c1; };                  // c0 + c1 w, w^2 = v
inline Fq12 add(const Fq12& a, const Fq12& b) { return {add(a.c0, b.c0), add(a.c1, b.c1)}; }
inline Fq12 sub(const Fq12& a, const Fq12& b) { return {sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
inline bool eq(const Fq12& a, const Fq12& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
inline Fq12 mul(const Fq12& a, const Fq12& b) {
    const Fq6 t0 = mul(a.c0, b.c0), t1 = mul(a.c1, b.c1);
    return {add(t0, mul_by_v(t1)), sub(sub(mul(add(a.c0, a.c1), add(b.c0, b.c1)), t0), t1)};
}
inline Fq12 conj(const Fq12& a) { return {a.c0, neg(a.c1)}; }      // the p^6 Frobenius
inline Fq12 inv(const Fq12& a) {
    const Fq6 d = inv(sub(mul(a.c0, a.c0), mul_by_v(mul(a.c1, a.c1))));
    return {mul(a.c0, d), neg(mul(a.c1, d))};
}
static const Fq12 kOne12 = {kOne6, kZero6};
inline Fq12 from_fq(const Fq& a) { return {{{a, kZero}, kZero2, kZero2}, kZero6}; }
inline Fq12 from_fq2(const Fq2& a) { return {{a, kZero2, kZero2}, kZero6}; }

struct G1Affine { Fq x, y; bool infinity; };
struct G2Affine { Fq2 x, y; bool infinity; };

// ---- the groups, affine (what a verifier needs: a handful of additions and two scalar multiplications per opening) -------------------------
// the generators ark-ec uses (canonical limbs)
static const uint64_t kG1x[6] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull, 0xc3688c4f9774b905ull, 0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull}, kG1y[6] = {0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull, 0x00db18cb2c04b3edull, 0xfcf5e095d5d00af6ull, 0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};
static const uint64_t kG2x0[6] = {0xd48056c8c121bdb8ull, 0x0bac0326a805bbefull, 0xb4510b647ae3d177ull, 0xc6e47ad4fa403b02ull, 0x260805272dc51051ull, 0x024aa2b2f08f0a91ull}, kG2x1[6] = {0xe5ac7d055d042b7eull, 0x334cf11213945d57ull, 0xb5da61bbdc7f5049ull, 0x596bd0d09920b61aull, 0x7dacd3a088274f65ull, 0x13e02b6052719f60ull}, kG2y0[6] = {0xe193548608b82801ull, 0x923ac9cc3baca289ull, 0x6d429a695160d12cull, 0xadfd9baa8cbdd3a7ull, 0x8cc9cdc6da2e351aull, 0x0ce5d527727d6e11ull}, kG2y1[6] = {0xaaa9075ff05f79beull, 0x3f370d275cec1da1ull, 0x267492ab572e99abull, 0xcb3e287e85a763afull, 0x32acd2b02bc28b99ull, 0x0606c4a02ea734ccull};
inline G1Affine g1_generator() { return {from_canonical(kG1x), from_canonical(kG1y), false}; }
inline G2Affine g2_generator() { return {{from_canonical(kG2x0), from_canonical(kG2x1)}, {from_canonical(kG2y0), from_canonical(kG2y1)}, false}; }
inline Fq mul3(const Fq& a) { return add(add(a, a), a); }
inline Fq2 mul3(const Fq2& a) { return add(add(a, a), a); }
template <class P, class F>
inline P ec_add(const P& a, const P& b) {
    if (a.infinity) return b;
    if (b.infinity) return a;
    F lam;
    if (eq(a.x, b.x)) {
        if (is_zero(add(a.y, b.y))) { P o = a; o.infinity = true; return o; }
        lam = mul(mul3(mul(a.x, a.x)), inv(add(a.y, a.y)));
    } else {
        lam = mul(sub(b.y, a.y), inv(sub(b.x, a.x)));
    }
    P r;
    r.infinity = false;
    r.x = sub(sub(mul(lam, lam), a.x), b.x);
    r.y = sub(mul(lam, sub(a.x, r.x)), a.y);
    return r;
}
inline G1Affine g1_add(const G1Affine& a, const G1Affine& b) { return ec_add<G1Affine, Fq>(a, b); }
inline G2Affine g2_add(const G2Affine& a, const G2Affine& b) { return ec_add<G2Affine, Fq2>(a, b); }
inline G1Affine g1_neg(const G1Affine& a) { return {a.x, neg(a.y), a.infinity}; }
inline G2Affine g2_neg(const G2Affine& a) { return {a.x, neg(a.y), a.infinity}; }
// k * P, k as four canonical 64-bit limbs (a scalar-field element)
template <class P, class F>
inline P ec_mul(const uint64_t k[4], P pt) {
    P acc = pt;
    acc.infinity = true;
    for (int i = 0; i < 256; i++) {
        if ((k[i / 64] >> (i % 64)) & 1) acc = ec_add<P, F>(acc, pt);
        pt = ec_add<P, F>(pt, pt);
    }
    return acc;
}
inline G1Affine g1_mul(const uint64_t k[4], const G1Affine& p) { return ec_mul<G1Affine, Fq>(k, p); }
inline G2Affine g2_mul(const uint64_t k[4], const G2Affine& p) { return ec_mul<G2Affine, Fq2>(k, p); }
// ark-ec's G1Projective in memory: Jacobian (X, Y, Z), each coordinate 6 Montgomery limbs (the same R = 2^384): x = X / Z^2, y = Y / Z^3
inline bool g1_from_ark(const uint64_t* a, G1Affine* out) {
    for (int c = 0; c < 3; c++)
        if (geq_p(a + 6 * c)) return false;
    Fq X, Y, Z;
    memcpy(X.v, a, 48); memcpy(Y.v, a + 6, 48); memcpy(Z.v, a + 12, 48);
    if (is_zero(Z)) { out->infinity = true; out->x = kZero; out->y = kZero; return true; }
    const Fq zi = inv(Z), zi2 = mul(zi, zi);
    out->infinity = false;
    out->x = mul(X, zi2);
    out->y = mul(Y, mul(zi2, zi));
    return eq(mul(out->y, out->y), add(mul(mul(out->x, out->x), out->x), from_u64(4)));
}

// f_{|x|, psi(Q)}(P), conjugated for x < 0; 1 when either point is the identity
inline Fq12 miller_loop(const G1Affine& p, const G2Affine& q) {
    if (p.infinity || q.infinity) return kOne12;
    static const Fq12 w = {kZero6, kOne6};
    static const Fq12 w2_inv = inv(mul(w, w)), w3_inv = inv(mul(mul(w, w), w));
    const Fq12 xp = from_fq(p.x), yp = from_fq(p.y);
    const Fq12 qx = mul(from_fq2(q.x), w2_inv), qy = mul(from_fq2(q.y), w3_inv);      // psi(Q) on E(Fq12): y^2 = x^3 + 4
    Fq12 tx = qx, ty = qy, f = kOne12;
    const Fq12 three = from_fq(from_u64(3));
    for (int i = 62; i >= 0; i--) {           // the bits of |x| below its leading one (bit 63)
        Fq12 lam = mul(mul(three, mul(tx, tx)), inv(add(ty, ty)));                     // tangent at T
        f = mul(mul(f, f), sub(sub(yp, ty), mul(lam, sub(xp, tx))));
        Fq12 nx = sub(sub(mul(lam, lam), tx), tx);
        ty = sub(mul(lam, sub(tx, nx)), ty);
        tx = nx;
        if ((kXAbs >> i) & 1) {
            lam = mul(sub(qy, ty), inv(sub(qx, tx)));                                  // chord through T and Q
            f = mul(f, sub(sub(yp, ty), mul(lam, sub(xp, tx))));
            nx = sub(sub(mul(lam, lam), tx), qx);
            ty = sub(mul(lam, sub(tx, nx)), ty);
            tx = nx;
        }
    }
    return conj(f);
}
inline Fq12 final_exponentiation(const Fq12& f) {
    Fq12 acc = kOne12;
    bool started = false;
    for (int i = 68 * 64 - 1; i >= 0; i--) {
        const bool bit = (kFinalExp[i / 64] >> (i % 64)) & 1;
        if (started) acc = mul(acc, acc);
        if (bit) { acc = started ? mul(acc, f) : f; started = true; }
    }
    return acc;
}

}  // namespace pairing
}  // namespace zksc
