// BLS12-381 scalar field (Fr) arithmetic for sm_100a -- 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// Replaces the arithmetic the reference gets from ark-ff 0.4.2 `Fp256<MontBackend<FrConfig,4>>`
// (third-party crate, not under /root/reference; call sites e.g.
// polynomial/src/multilinear/evaluation_form.rs:133, polynomial/src/composed/composed_multilinear.rs:109).
// The in-memory element is bit-identical to ark-ff's: 4 x u64 little-endian limbs == 8 x u32
// little-endian limbs, Montgomery form, 32 bytes, so a `&[Fr]` can be copied to the device verbatim.
//
// Design notes (see DESIGN.md "Kernels"):
//  * all multi-limb products are carry chains of mad.lo.cc / madc.hi.cc pairs; ptxas fuses each pair
//    into one IMAD.WIDE.U32(.X) with a predicate carry (checked with cuobjdump -sass);
//  * the 512-bit product is built "even/odd": for one multiplier limb b_i the products with
//    a0,a2,a4,a6 form one clean carry chain and those with a1,a3,a5,a7 a second one, so no
//    instruction is spent on per-product carry fix-up;
//  * the modulus is r = 1 + 2^32*q, so -r^-1 mod 2^32 = 0xffffffff and the Montgomery quotient
//    limb is just m = -T[i];
//  * sums of products are accumulated UNREDUCED (17 limbs) and reduced once per block.
#pragma once
#include <cstdint>
#ifndef ZKSC_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace zksc {

struct alignas(32) Fr {
    uint32_t l[8];
};

// r, little-endian 32-bit limbs
#define ZKSC_P0 0x00000001u
#define ZKSC_P1 0xffffffffu
#define ZKSC_P2 0xfffe5bfeu
#define ZKSC_P3 0x53bda402u
#define ZKSC_P4 0x09a1d805u
#define ZKSC_P5 0x3339d808u
#define ZKSC_P6 0x299d7d48u
#define ZKSC_P7 0x73eda753u

namespace ptx {
#ifndef ZKSC_HOST_EMU
#define ZKSC_DEV __device__ __forceinline__
ZKSC_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKSC_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKSC_DEV uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKSC_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKSC_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKSC_DEV uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKSC_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKSC_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKSC_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKSC_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKSC_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKSC_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation of the PTX extended-precision instructions (tests/emu only): CC.CF is a thread-local.
#define ZKSC_DEV static inline
static thread_local uint32_t CF = 0;
ZKSC_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKSC_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + CF; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKSC_DEV uint32_t addc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a + b + CF); }
ZKSC_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; CF = (uint32_t)(s >> 63); return (uint32_t)s; }
ZKSC_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - CF; CF = (uint32_t)(s >> 63); return (uint32_t)s; }
ZKSC_DEV uint32_t subc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a - b - CF); }
ZKSC_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
ZKSC_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
ZKSC_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_lo(a, b) + c; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKSC_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_lo(a, b) + c + CF; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKSC_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_hi(a, b) + c + CF; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
ZKSC_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((uint64_t)mul_hi(a, b) + c + CF); }
#endif
}  // namespace ptx

#ifndef ZKSC_HOST_EMU
// ---- 256-bit vector memory access (LDG.E.256 / STG.E.256 on sm_100a) --------------------------
ZKSC_DEV Fr ld256(const Fr* p) {
    Fr v;
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]), "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7])
                 : "l"(p));
    return v;
}
// read-only, streaming (no L1 allocation): for tables that are never written by the running kernel
ZKSC_DEV Fr ld256_stream(const Fr* p) {
    Fr v;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]), "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7])
                 : "l"(p));
    return v;
}
ZKSC_DEV void st256(Fr* p, const Fr& v) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.l[0]), "r"(v.l[1]), "r"(v.l[2]), "r"(v.l[3]),
                 "r"(v.l[4]), "r"(v.l[5]), "r"(v.l[6]), "r"(v.l[7])
                 : "memory");
}

#endif  // !ZKSC_HOST_EMU

ZKSC_DEV Fr fr_zero() {
    Fr z;
#pragma unroll
    for (int i = 0; i < 8; i++) z.l[i] = 0;
    return z;
}
// Montgomery form of 1 (= 2^256 mod r)
ZKSC_DEV Fr fr_one() {
    Fr o;
    o.l[0] = 0xfffffffeu; o.l[1] = 0x00000001u; o.l[2] = 0x00034802u; o.l[3] = 0x5884b7fau;
    o.l[4] = 0xecbc4ff5u; o.l[5] = 0x998c4fefu; o.l[6] = 0xacc5056fu; o.l[7] = 0x1824b159u;
    return o;
}
// R^2 mod r (raw limbs): mont_mul(x, R2) converts canonical x to Montgomery form
ZKSC_DEV Fr fr_r2() {
    Fr o;
    o.l[0] = 0xf3f29c6du; o.l[1] = 0xc999e990u; o.l[2] = 0x87925c23u; o.l[3] = 0x2b6cedcbu;
    o.l[4] = 0x7254398fu; o.l[5] = 0x05d31496u; o.l[6] = 0x9f59ff11u; o.l[7] = 0x0748d9d9u;
    return o;
}

// x (< 2^256) -> x - r if x >= r.  One conditional subtraction.
ZKSC_DEV void cond_sub_r(uint32_t (&x)[8]) {
    uint32_t t[8];
    t[0] = ptx::sub_cc(x[0], ZKSC_P0);
    t[1] = ptx::subc_cc(x[1], ZKSC_P1);
    t[2] = ptx::subc_cc(x[2], ZKSC_P2);
    t[3] = ptx::subc_cc(x[3], ZKSC_P3);
    t[4] = ptx::subc_cc(x[4], ZKSC_P4);
    t[5] = ptx::subc_cc(x[5], ZKSC_P5);
    t[6] = ptx::subc_cc(x[6], ZKSC_P6);
    t[7] = ptx::subc_cc(x[7], ZKSC_P7);
    uint32_t borrow = ptx::subc(0u, 0u);  // 0xffffffff if x < r
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : t[i];
}

ZKSC_DEV Fr fr_add(const Fr& a, const Fr& b) {  // a,b < r  ->  (a+b) mod r
    Fr s;
    s.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) s.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    s.l[7] = ptx::addc(a.l[7], b.l[7]);  // a+b < 2r < 2^256: no carry out
    cond_sub_r(s.l);
    return s;
}

ZKSC_DEV Fr fr_sub(const Fr& a, const Fr& b) {  // a,b < r  ->  (a-b) mod r
    Fr d;
    d.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) d.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
    uint32_t mask = ptx::subc(0u, 0u);  // all-ones if a < b
    d.l[0] = ptx::add_cc(d.l[0], mask & ZKSC_P0);
    d.l[1] = ptx::addc_cc(d.l[1], mask & ZKSC_P1);
    d.l[2] = ptx::addc_cc(d.l[2], mask & ZKSC_P2);
    d.l[3] = ptx::addc_cc(d.l[3], mask & ZKSC_P3);
    d.l[4] = ptx::addc_cc(d.l[4], mask & ZKSC_P4);
    d.l[5] = ptx::addc_cc(d.l[5], mask & ZKSC_P5);
    d.l[6] = ptx::addc_cc(d.l[6], mask & ZKSC_P6);
    d.l[7] = ptx::addc(d.l[7], mask & ZKSC_P7);
    return d;
}

// x < 2^256 (any residue representative)  ->  canonical x mod r   (2^256 / r = 2.2: two subtractions)
ZKSC_DEV Fr fr_canon(const Fr& x) {
    Fr y = x;
    cond_sub_r(y.l);
    cond_sub_r(y.l);
    return y;
}

// ---- multi-limb products --------------------------------------------------------------------------
// IMAD.WIDE needs its 64-bit accumulator in an aligned register pair, so partial products whose low
// limb sits at an EVEN position accumulate in E[] and those at an ODD position in O[] (index =
// absolute limb position).  For one multiplier limb x at base position i the products with
// a0,a2,a4,a6 form one carry chain in the array of i's parity and those with a1,a3,a5,a7 a second
// chain in the other array; the carry out of a chain lands on a limb that is still fresh (small).
#ifndef ZKSC_HOST_EMU
static __device__ __constant__ uint32_t kModulus[8] = {ZKSC_P0, ZKSC_P1, ZKSC_P2, ZKSC_P3, ZKSC_P4, ZKSC_P5, ZKSC_P6, ZKSC_P7};
#else
static const uint32_t kModulus[8] = {ZKSC_P0, ZKSC_P1, ZKSC_P2, ZKSC_P3, ZKSC_P4, ZKSC_P5, ZKSC_P6, ZKSC_P7};
#endif

// The modulus limbs are fetched with a volatile ld.const so that ptxas sees opaque values: with
// immediates (or a constant bank whose contents it knows) it strength-reduces the multiplications by
// 1 and 0xffffffff and no longer fuses the lo/hi pair into one IMAD.WIDE; as opaque values they live
// in uniform registers (no vector registers spent) and every pair fuses (checked in SASS).
ZKSC_DEV void load_modulus(uint32_t& m0, uint32_t& m1, uint32_t& m2, uint32_t& m3, uint32_t& m4, uint32_t& m5, uint32_t& m6, uint32_t& m7) {
#ifndef ZKSC_HOST_EMU
    // &kModulus is a GENERIC address in device code: convert it to a .const-space address first
    // (an ld.const on the generic address reads garbage).
    asm volatile("{\n\t.reg .u64 cp;\n\tcvta.to.const.u64 cp, %8;\n\t"
                 "ld.const.u32 %0, [cp]; ld.const.u32 %1, [cp+4]; ld.const.u32 %2, [cp+8]; ld.const.u32 %3, [cp+12];\n\t"
                 "ld.const.u32 %4, [cp+16]; ld.const.u32 %5, [cp+20]; ld.const.u32 %6, [cp+24]; ld.const.u32 %7, [cp+28];\n\t}"
                 : "=r"(m0), "=r"(m1), "=r"(m2), "=r"(m3), "=r"(m4), "=r"(m5), "=r"(m6), "=r"(m7)
                 : "l"(kModulus));
#else
    m0 = kModulus[0]; m1 = kModulus[1]; m2 = kModulus[2]; m3 = kModulus[3];
    m4 = kModulus[4]; m5 = kModulus[5]; m6 = kModulus[6]; m7 = kModulus[7];
#endif
}

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

// Montgomery product, canonical result.  a, b < r.
ZKSC_DEV Fr fr_mul(const Fr& a, const Fr& b) {
    Fr o;
    uint32_t top;
    mont_mul_raw(o.l, top, a, b);   // < 2r < 2^256: top == 0
    cond_sub_r(o.l);
    return o;
}

// Montgomery reduction of a 16-limb T (any T < 2^512): (T + M r) / 2^256, 8 limbs + top.
// Only used on the once-per-block slow path and in tests.
ZKSC_DEV void redc_rows(uint32_t (&T)[16], uint32_t& top) {
    using namespace ptx;
    top = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t m = 0u - T[i];
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            // T[i+j] += lo, T[i+j+1] += hi  (plain 64-bit accumulation; clarity over speed)
            unsigned long long w = (unsigned long long)m * kModulus[j] + T[i + j] + c;
            T[i + j] = (uint32_t)w;
            c = (uint32_t)(w >> 32);
        }
#pragma unroll
        for (int p = i + 8; p < 16; p++) {
            unsigned long long w = (unsigned long long)T[p] + c;
            T[p] = (uint32_t)w;
            c = (uint32_t)(w >> 32);
        }
        top += c;
    }
}

// ---- wide (unreduced) accumulators -------------------------------------------------------------
template <int NL>
struct Acc {
    uint32_t l[NL];
};
template <int NL>
ZKSC_DEV void acc_zero(Acc<NL>& a) {
#pragma unroll
    for (int i = 0; i < NL; i++) a.l[i] = 0;
}
// acc += x, x has NX <= NL limbs
template <int NL, int NX>
ZKSC_DEV void acc_add(Acc<NL>& a, const uint32_t (&x)[NX]) {
    static_assert(NX <= NL, "");
    a.l[0] = ptx::add_cc(a.l[0], x[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) {
        const uint32_t xi = (i < NX) ? x[i] : 0u;
        if (i < NL - 1) a.l[i] = ptx::addc_cc(a.l[i], xi);
        else a.l[i] = ptx::addc(a.l[i], xi);
    }
}
template <int NL>
ZKSC_DEV void acc_add_acc(Acc<NL>& a, const Acc<NL>& b) {
    acc_add<NL, NL>(a, b.l);
}
#ifndef ZKSC_HOST_EMU
template <int NL>
ZKSC_DEV Acc<NL> acc_shfl_down(const Acc<NL>& a, int delta) {
    Acc<NL> o;
#pragma unroll
    for (int i = 0; i < NL; i++) o.l[i] = __shfl_down_sync(0xffffffffu, a.l[i], delta);
    return o;
}
template <int NL>
ZKSC_DEV void acc_warp_reduce(Acc<NL>& a) {  // lane 0 ends up with the warp total
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        Acc<NL> o = acc_shfl_down(a, d);
        acc_add_acc(a, o);
    }
}

#endif  // !ZKSC_HOST_EMU

// 9-limb sum of Montgomery-form elements  ->  canonical Montgomery element of the sum.
// A = L + h * 2^256  =>  A mod r = canon(L) + h * (2^256 mod r) = canon(L) + mont_mul(h, R^2)
ZKSC_DEV Fr acc9_reduce(const Acc<9>& a) {
    Fr lo, h = fr_zero();
#pragma unroll
    for (int i = 0; i < 8; i++) lo.l[i] = a.l[i];
    h.l[0] = a.l[8];
    return fr_add(fr_canon(lo), fr_mul(h, fr_r2()));
}
// 17-limb sum of unreduced Montgomery products (each a*b, a and b Montgomery form)  ->  canonical
// Montgomery element of sum(a*b / R).   A = L + M * 2^256 + h * 2^512:
//   A / R mod r = L * R^-1 + M + h * R = mont_mul(canon(L), 1) + canon(M) + mont_mul(h, R^2)
ZKSC_DEV Fr acc17_reduce(const Acc<17>& a) {
    Fr lo, mid, h = fr_zero(), one = fr_zero();
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.l[i] = a.l[i]; mid.l[i] = a.l[8 + i]; }
    h.l[0] = a.l[16];
    one.l[0] = 1u;
    Fr r0 = fr_mul(fr_canon(lo), one);
    Fr r1 = fr_canon(mid);
    Fr r2 = fr_mul(h, fr_r2());
    return fr_add(fr_add(r0, r1), r2);
}

// fold: a + r * (b - a)   (== r*b + (1-r)*a, polynomial/src/multilinear/evaluation_form.rs:133)
ZKSC_DEV Fr fr_fold(const Fr& a, const Fr& b, const Fr& r) {
    return fr_add(a, fr_mul(r, fr_sub(b, a)));
}

}  // namespace zksc
