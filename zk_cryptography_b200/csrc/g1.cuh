// BLS12-381 G1 multi-scalar multiplication for sm_100a -- the device half of MultilinearKZG::commitment / open
// (kzg/src/multilinear_kzg.rs:33-88 of the reference: sum_i  evaluations[i] * powers_of_tau_in_g1[i], which the reference computes
// with one `mul_bigint` per term), SURVEY 8(f) next-4.
//
// The curve arithmetic the reference gets from ark-ec / ark-test-curves 0.4.2 (third-party, not under /root/reference):
//   base field Fq, 381 bits, 12 x 32-bit limbs, Montgomery form with R = 2^384 (ark-ff's in-memory form: 6 x u64 little-endian);
//   G1: y^2 = x^3 + 4, Jacobian coordinates (X, Y, Z), x = X / Z^2, y = Y / Z^3, Z = 0 <=> infinity -- the memory layout of ark-ec's
//   short_weierstrass::Projective, so a `&[G1Projective]` crosses the C ABI as 18 x u64 per point without conversion.
//
// Method (Pippenger buckets, no sorting, no atomics): the 255-bit scalars are cut into 32 windows of 8 bits; one THREAD owns one
// (window, bucket) pair and scans the window's digit row, adding every point whose digit is its bucket number -- for scalars
// that are uniformly distributed every thread adds ~n / 256 points, and the 32 lanes of a warp read the same digit at the same
// time (one broadcast load).  Then one thread per window folds its 255 buckets with the running-sum trick, one thread combines the
// windows, and the result is normalised (Z = 1).  This is a first, parity-checked slice: correct and parallel, not yet tuned.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace zksc {
namespace g1 {

constexpr int kL = 12;      // 32-bit limbs of an Fq element
struct Fq {
    uint32_t l[kL];
};
struct Jac {      // 144 bytes = ark-ec Projective<g1::Config>: x, y, z, each 6 x u64 little-endian Montgomery
    Fq x, y, z;
};

__device__ __constant__ const uint32_t kP[kL] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                                 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
__device__ __constant__ const uint32_t kOneMont[kL] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                                       0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};   // 2^384 mod p
constexpr uint32_t kN0 = 0xfffcfffdu;     // -p^-1 mod 2^32

__device__ __forceinline__ Fq fq_zero() {
    Fq z;
#pragma unroll
    for (int i = 0; i < kL; i++) z.l[i] = 0;
    return z;
}
__device__ __forceinline__ Fq fq_one() {
    Fq o;
#pragma unroll
    for (int i = 0; i < kL; i++) o.l[i] = kOneMont[i];
    return o;
}
__device__ __forceinline__ bool fq_is_zero(const Fq& a) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < kL; i++) x |= a.l[i];
    return x == 0;
}
__device__ __forceinline__ bool fq_eq(const Fq& a, const Fq& b) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < kL; i++) x |= a.l[i] ^ b.l[i];
    return x == 0;
}
// a >= p ?
__device__ __forceinline__ bool fq_geq_p(const Fq& a) {
#pragma unroll
    for (int i = kL - 1; i >= 0; i--) {
        if (a.l[i] != kP[i]) return a.l[i] > kP[i];
    }
    return true;
}
__device__ __forceinline__ void fq_sub_p(Fq& a) {
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < kL; i++) {
        const uint64_t d = (uint64_t)a.l[i] - kP[i] - br;
        a.l[i] = (uint32_t)d;
        br = (d >> 63) & 1;
    }
}
__device__ __forceinline__ Fq fq_add(const Fq& a, const Fq& b) {      // a, b < p; 2p < 2^384: no carry out
    Fq s;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < kL; i++) {
        const uint64_t t = (uint64_t)a.l[i] + b.l[i] + c;
        s.l[i] = (uint32_t)t;
        c = t >> 32;
    }
    if (fq_geq_p(s)) fq_sub_p(s);
    return s;
}
__device__ __forceinline__ Fq fq_sub(const Fq& a, const Fq& b) {
    Fq d;
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < kL; i++) {
        const uint64_t t = (uint64_t)a.l[i] - b.l[i] - br;
        d.l[i] = (uint32_t)t;
        br = (t >> 63) & 1;
    }
    if (br) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < kL; i++) {
            const uint64_t t = (uint64_t)d.l[i] + kP[i] + c;
            d.l[i] = (uint32_t)t;
            c = t >> 32;
        }
    }
    return d;
}
__device__ __forceinline__ Fq fq_dbl(const Fq& a) { return fq_add(a, a); }
// Montgomery product (CIOS, 32-bit limbs, 64-bit accumulation): a b / 2^384 mod p, canonical
__device__ __noinline__ Fq fq_mul(const Fq& a, const Fq& b) {
    uint32_t t[kL + 2];
#pragma unroll
    for (int i = 0; i < kL + 2; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < kL; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < kL; j++) {
            const uint64_t s = (uint64_t)a.l[j] * b.l[i] + t[j] + c;
            t[j] = (uint32_t)s;
            c = s >> 32;
        }
        uint64_t s = (uint64_t)t[kL] + c;
        t[kL] = (uint32_t)s;
        t[kL + 1] = (uint32_t)(s >> 32);
        const uint32_t m = t[0] * kN0;
        c = ((uint64_t)m * kP[0] + t[0]) >> 32;
#pragma unroll
        for (int j = 1; j < kL; j++) {
            const uint64_t u = (uint64_t)m * kP[j] + t[j] + c;
            t[j - 1] = (uint32_t)u;
            c = u >> 32;
        }
        s = (uint64_t)t[kL] + c;
        t[kL - 1] = (uint32_t)s;
        t[kL] = t[kL + 1] + (uint32_t)(s >> 32);
    }
    Fq r;
#pragma unroll
    for (int i = 0; i < kL; i++) r.l[i] = t[i];
    if (t[kL] || fq_geq_p(r)) fq_sub_p(r);
    return r;
}
__device__ __forceinline__ Fq fq_sqr(const Fq& a) { return fq_mul(a, a); }
// a^(p-2): the inverse (one per MSM, for the final normalisation)
__device__ __noinline__ Fq fq_inv(const Fq& a) {
    // exponent p - 2, scanned from the top bit
    Fq acc = fq_one();
    for (int i = kL * 32 - 1; i >= 0; i--) {
        acc = fq_sqr(acc);
        uint32_t limb = kP[i >> 5];
        if ((i >> 5) == 0) limb -= 2;            // p - 2: the low limb 0xffffaaab - 2 does not borrow
        if ((limb >> (i & 31)) & 1) acc = fq_mul(acc, a);
    }
    return acc;
}

__device__ __forceinline__ Jac jac_infinity() {
    Jac p;
    p.x = fq_one(); p.y = fq_one(); p.z = fq_zero();      // ark-ec's representation of the identity
    return p;
}
__device__ __forceinline__ bool jac_is_inf(const Jac& p) { return fq_is_zero(p.z); }
// dbl-2009-l (a = 0)
__device__ __noinline__ Jac jac_double(const Jac& p) {
    if (jac_is_inf(p)) return p;
    const Fq A = fq_sqr(p.x), B = fq_sqr(p.y), C = fq_sqr(B);
    Fq D = fq_sub(fq_sub(fq_sqr(fq_add(p.x, B)), A), C);
    D = fq_dbl(D);
    const Fq E = fq_add(fq_dbl(A), A), F = fq_sqr(E);
    Jac r;
    r.x = fq_sub(F, fq_dbl(D));
    Fq C8 = fq_dbl(fq_dbl(fq_dbl(C)));
    r.y = fq_sub(fq_mul(E, fq_sub(D, r.x)), C8);
    r.z = fq_dbl(fq_mul(p.y, p.z));
    return r;
}
// add-2007-bl with the exceptional cases handled (either operand at infinity, equal points, opposite points)
__device__ __noinline__ Jac jac_add(const Jac& p, const Jac& q) {
    if (jac_is_inf(p)) return q;
    if (jac_is_inf(q)) return p;
    const Fq Z1Z1 = fq_sqr(p.z), Z2Z2 = fq_sqr(q.z);
    const Fq U1 = fq_mul(p.x, Z2Z2), U2 = fq_mul(q.x, Z1Z1);
    const Fq S1 = fq_mul(fq_mul(p.y, q.z), Z2Z2), S2 = fq_mul(fq_mul(q.y, p.z), Z1Z1);
    if (fq_eq(U1, U2)) {
        if (fq_eq(S1, S2)) return jac_double(p);
        return jac_infinity();
    }
    const Fq H = fq_sub(U2, U1), I = fq_sqr(fq_dbl(H)), J = fq_mul(H, I);
    const Fq rr = fq_dbl(fq_sub(S2, S1)), V = fq_mul(U1, I);
    Jac r;
    r.x = fq_sub(fq_sub(fq_sqr(rr), J), fq_dbl(V));
    r.y = fq_sub(fq_mul(rr, fq_sub(V, r.x)), fq_dbl(fq_mul(S1, J)));
    r.z = fq_mul(fq_sub(fq_sub(fq_sqr(fq_add(p.z, q.z)), Z1Z1), Z2Z2), H);
    return r;
}

constexpr int kWindowBits = 8, kWindows = 32, kBuckets = 255;      // 32 x 8 = 256 >= 255 scalar bits

// digits[w][i] = bits 8w .. 8w+7 of the canonical value of scalar i (scalars arrive as Montgomery-form Fr: ark-ff's memory form)
__global__ void __launch_bounds__(256) msm_digits_kernel(const Fr* scalars, unsigned long long n, unsigned long long period, uint8_t* digits) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    Fr one = fr_zero();
    one.l[0] = 1u;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const Fr c = fr_mul(ld256(scalars + (period ? i % period : i)), one);      // Montgomery -> canonical
#pragma unroll
        for (int w = 0; w < kWindows; w++) digits[(size_t)w * n + i] = (uint8_t)(c.l[w >> 2] >> (8 * (w & 3)));
    }
}
// one thread per (window, bucket): buckets[w][b - 1] = sum of the points whose window-w digit is b
__global__ void __launch_bounds__(128) msm_bucket_kernel(const uint8_t* digits, const Jac* points, unsigned long long n, Jac* buckets) {
    const unsigned int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= kWindows * 256) return;
    const unsigned int w = id >> 8, b = id & 255;
    if (b == 0) return;
    const uint8_t* row = digits + (size_t)w * n;
    Jac acc = jac_infinity();
    for (unsigned long long i = 0; i < n; i++)
        if (row[i] == b) acc = jac_add(acc, points[i]);
    buckets[(size_t)w * kBuckets + (b - 1)] = acc;
}
// one thread per window: sum_b b * bucket[b] by running sums (from the top bucket down)
__global__ void msm_window_kernel(const Jac* buckets, Jac* windows) {
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= kWindows) return;
    Jac running = jac_infinity(), total = jac_infinity();
    for (int b = kBuckets - 1; b >= 0; b--) {
        running = jac_add(running, buckets[(size_t)w * kBuckets + b]);
        total = jac_add(total, running);
    }
    windows[w] = total;
}
// one thread: sum_w 2^(8w) windows[w], normalised to Z = 1 (or the identity)
__global__ void msm_combine_kernel(const Jac* windows, Jac* out) {
    if (blockIdx.x || threadIdx.x) return;
    Jac acc = jac_infinity();
    for (int w = kWindows - 1; w >= 0; w--) {
        for (int k = 0; k < kWindowBits; k++) acc = jac_double(acc);
        acc = jac_add(acc, windows[w]);
    }
    if (!jac_is_inf(acc)) {
        const Fq zi = fq_inv(acc.z), zi2 = fq_sqr(zi);
        acc.x = fq_mul(acc.x, zi2);
        acc.y = fq_mul(acc.y, fq_mul(zi2, zi));
        acc.z = fq_one();
    } else {
        acc = jac_infinity();
    }
    *out = acc;
}

}  // namespace g1
}  // namespace zksc
