// EVIDENCE ONLY (tools/ubench*.cu): the first-generation row-wise (operand scanning, even/odd carry chain)
// multiplier, kept so the microbenchmarks can show what the product-scanning multiplier in
// zk_cryptography_b200/csrc/fr.cuh replaced.  Not included by the library.
#pragma once
#include "../zk_cryptography_b200/csrc/fr.cuh"
namespace zksc {
namespace rowwise {
static __device__ __constant__ uint32_t kModulus[8] = {ZKSC_P0, ZKSC_P1, ZKSC_P2, ZKSC_P3, ZKSC_P4, ZKSC_P5, ZKSC_P6, ZKSC_P7};
ZKSC_DEV void load_modulus(uint32_t& m0, uint32_t& m1, uint32_t& m2, uint32_t& m3, uint32_t& m4, uint32_t& m5, uint32_t& m6, uint32_t& m7) {
    asm volatile("{\n\t.reg .u64 cp;\n\tcvta.to.const.u64 cp, %8;\n\t"
                 "ld.const.u32 %0, [cp]; ld.const.u32 %1, [cp+4]; ld.const.u32 %2, [cp+8]; ld.const.u32 %3, [cp+12];\n\t"
                 "ld.const.u32 %4, [cp+16]; ld.const.u32 %5, [cp+20]; ld.const.u32 %6, [cp+24]; ld.const.u32 %7, [cp+28];\n\t}"
                 : "=r"(m0), "=r"(m1), "=r"(m2), "=r"(m3), "=r"(m4), "=r"(m5), "=r"(m6), "=r"(m7)
                 : "l"(kModulus));
}
// ---- multi-limb products --------------------------------------------------------------------------
// IMAD.WIDE needs its 64-bit accumulator in an aligned register pair, so partial products whose low
// limb sits at an EVEN position accumulate in E[] and those at an ODD position in O[] (index =
// absolute limb position).  For one multiplier limb x at base position i the products with
// a0,a2,a4,a6 form one carry chain in the array of i's parity and those with a1,a3,a5,a7 a second
// chain in the other array; the carry out of a chain lands on a limb that is still fresh (small).
// X[pos .. pos+7] += (v0, v2, v4, v6 as limbs 0,2,4,6) * x ; carry into X[pos+8].
// CARRY_IN: the chain starts with the pending CC.CF.
template <bool CARRY_IN>
ZKSC_DEV void chain4(uint32_t* X, int pos, uint32_t v0, uint32_t v2, uint32_t v4, uint32_t v6, uint32_t x) {
    using namespace ptx;
    X[pos + 0] = CARRY_IN ? madc_lo_cc(x, v0, X[pos + 0]) : mad_lo_cc(x, v0, X[pos + 0]);
    X[pos + 1] = madc_hi_cc(x, v0, X[pos + 1]);
    X[pos + 2] = madc_lo_cc(x, v2, X[pos + 2]); X[pos + 3] = madc_hi_cc(x, v2, X[pos + 3]);
    X[pos + 4] = madc_lo_cc(x, v4, X[pos + 4]); X[pos + 5] = madc_hi_cc(x, v4, X[pos + 5]);
    X[pos + 6] = madc_lo_cc(x, v6, X[pos + 6]); X[pos + 7] = madc_hi_cc(x, v6, X[pos + 7]);
    X[pos + 8] = addc(X[pos + 8], 0u);
}

// 512-bit product a*b, any a, b < 2^256, as E + O (O[p] has weight 2^(32p), O[0] unused = 0).
ZKSC_DEV void mul_wide_eo(uint32_t (&E)[17], uint32_t (&O)[17], const Fr& a, const Fr& b) {
#pragma unroll
    for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* A = (i & 1) ? O : E;
        uint32_t* B = (i & 1) ? E : O;
        chain4<false>(A, i, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
        chain4<false>(B, i + 1, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
    }
}
// T = E + (O) merged into 16 limbs (the product fits 512 bits)
ZKSC_DEV void merge_eo(uint32_t (&T)[16], const uint32_t (&E)[17], const uint32_t (&O)[17]) {
    T[0] = E[0];
    T[1] = ptx::add_cc(E[1], O[1]);
#pragma unroll
    for (int i = 2; i < 15; i++) T[i] = ptx::addc_cc(E[i], O[i]);
    T[15] = ptx::addc(E[15], O[15]);
}
ZKSC_DEV void mul_wide(uint32_t (&T)[16], const Fr& a, const Fr& b) {
    uint32_t E[17], O[17];
    mul_wide_eo(E, O, a, b);
    merge_eo(T, E, O);
}

// ---- Montgomery multiplication (CIOS, even/odd) --------------------------------------------------
// Row i adds a*b_i and then m_i*r at base position i, with m_i = -(limb i of the running total)
// because -r^-1 = -1 mod 2^32.  Limb i of the total is E[i] + O[i] + k, k being the carry produced
// when limb i-1 was cancelled; k enters the m*r chain as its carry-in.
// Inputs: a < 2^256 arbitrary, b < 2^256 arbitrary with a*b < r*2^256 for a result < 2r.
// Returns the (up to) 9-limb result a*b*2^-256 + (multiple of r), limbs 8..16 of the total.
ZKSC_DEV void mont_mul_raw(uint32_t (&res)[8], uint32_t& top, const Fr& a, const Fr& b) {
    using namespace ptx;
    uint32_t E[18], O[18];
#pragma unroll
    for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
    // The modulus limbs must sit in ordinary registers: with immediates or uniform registers ptxas
    // does not fuse the lo/hi pair into IMAD.WIDE (checked in SASS), which doubles the multiplies.
    uint32_t m0, m1, m2, m3, m4, m5, m6, m7;
    load_modulus(m0, m1, m2, m3, m4, m5, m6, m7);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* A = (i & 1) ? O : E;
        uint32_t* B = (i & 1) ? E : O;
        chain4<false>(A, i, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
        chain4<false>(B, i + 1, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
        if (i == 0) {
            const uint32_t m = 0u - E[0];
            chain4<false>(A, i, m0, m2, m4, m6, m);
            chain4<false>(B, i + 1, m1, m3, m5, m7, m);
        } else {
            (void)add_cc(E[i - 1], O[i - 1]);           // limb i-1 is 0 mod 2^32; CF = k_{i-1}
            const uint32_t t = addc(E[i], O[i]);
            const uint32_t m = 0u - t;
            chain4<true>(A, i, m0, m2, m4, m6, m);      // k_{i-1} enters here
            chain4<false>(B, i + 1, m1, m3, m5, m7, m);
        }
    }
    (void)add_cc(O[7], E[7]);                           // k_7
#pragma unroll
    for (int i = 0; i < 8; i++) res[i] = addc_cc(E[8 + i], O[8 + i]);
    top = addc(E[16], O[16]);
}


ZKSC_DEV Fr fr_mul(const Fr& a, const Fr& b) {
    Fr o;
    uint32_t top;
    mont_mul_raw(o.l, top, a, b);
    cond_sub_r(o.l);
    return o;
}
}  // namespace rowwise
}  // namespace zksc
