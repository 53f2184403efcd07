// BLS12-381 scalar field (Fr) arithmetic for sm_100a -- 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// Replaces the arithmetic the reference gets from ark-ff 0.4.2 `Fp256<MontBackend<FrConfig,4>>`
// (third-party crate, not under /root/reference; call sites e.g.
// polynomial/src/multilinear/evaluation_form.rs:133, polynomial/src/composed/composed_multilinear.rs:109).
// The in-memory element is bit-identical to ark-ff's: 4 x u64 little-endian limbs == 8 x u32
// little-endian limbs, Montgomery form, 32 bytes, so a `&[Fr]` can be copied to the device verbatim.
//
// Design notes (see DESIGN.md "Kernels"; measurements in profiles/r01_ubench*.jsonl):
//  * the multiplier pipe is the roof: IMAD.WIDE.U32 issues at ~31 per clock per SM on B200 (half the IMAD
//    rate), IMAD.HI.U32 at ~25, and FP64 DFMA shares the same pipe (no gain from mixing), so the aim is
//    ONE IMAD.WIDE per 32x32 limb product and nothing else on that pipe;
//  * multiplication is product scanning (column-wise) with a 96-bit accumulator: each partial product is
//    mad.lo.cc + madc.hi.cc, which ptxas fuses into one IMAD.WIDE.U32 with carry-out, plus an IADD3.X on
//    the ALU pipe that counts carries;
//  * the modulus is r = 1 + 2^32*u with r_1 = 2^32 - 1: the Montgomery digit needs no multiplication, the
//    digit * r_0 product is a carry, digit * r_1 is a shift and a subtraction (112 products in all), and
//    the digit is formed as a COMPLEMENT (~lo), never a negation -- see mul_ps;
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

// the same, out of line: for branches that are (almost) never taken, so that the hot path carries a compare and a branch only
// (by value: the element travels in registers, nothing is parked in local memory on the hot path)
#ifndef ZKSC_HOST_EMU
static __device__ __noinline__ Fr cond_sub_r_rare(Fr x) { cond_sub_r(x.l); return x; }
#else
static inline Fr cond_sub_r_rare(Fr x) { cond_sub_r(x.l); return x; }
#endif

ZKSC_DEV Fr fr_add(const Fr& a, const Fr& b) {  // a,b < r  ->  (a+b) mod r
    Fr s;
    s.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) s.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    s.l[7] = ptx::addc(a.l[7], b.l[7]);  // a+b < 2r < 2^256: no carry out
    cond_sub_r(s.l);
    return s;
}

// a, b < r  ->  a + b < 2r < 2^256, NOT reduced: for operands of mul_wide (any value < 2^256) and for ONE operand of
// fr_mul (a * b < r * 2^256 holds with the other operand < r)
ZKSC_DEV Fr fr_add_lazy(const Fr& a, const Fr& b) {
    Fr s;
    s.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) s.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    s.l[7] = ptx::addc(a.l[7], b.l[7]);
    return s;
}

// a, b < r  ->  a - b + r in (0, 2r), NOT reduced (no borrow test): for operands of mul_wide
ZKSC_DEV Fr fr_sub_lazy(const Fr& a, const Fr& b) {
    Fr d;
    d.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) d.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
    d.l[7] = ptx::subc(a.l[7], b.l[7]);
    d.l[0] = ptx::add_cc(d.l[0], ZKSC_P0); d.l[1] = ptx::addc_cc(d.l[1], ZKSC_P1); d.l[2] = ptx::addc_cc(d.l[2], ZKSC_P2);
    d.l[3] = ptx::addc_cc(d.l[3], ZKSC_P3); d.l[4] = ptx::addc_cc(d.l[4], ZKSC_P4); d.l[5] = ptx::addc_cc(d.l[5], ZKSC_P5);
    d.l[6] = ptx::addc_cc(d.l[6], ZKSC_P6); d.l[7] = ptx::addc(d.l[7], ZKSC_P7);
    return d;
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

// ---- product scanning (column-wise) multiplication ---------------------------------------------------
// One 96-bit column accumulator (lo, hi, cn).  Every partial product is ONE IMAD.WIDE.U32 with carry-OUT
// only (mad.lo.cc + madc.hi.cc, fused by ptxas) plus an IADD3.X that counts the carry (ptxas merges two
// carries into one IADD3.X); no IMAD.WIDE takes a carry-IN.
// REDUCE: Montgomery reduction interleaved (FIPS).  The textbook quotient digit is q_k = -lo_k (because
// -p^-1 = -1 mod 2^32), but ptxas folds a negation into the multiplier's operand modifiers and then
// splits every q*p product into IMAD + IMAD.HI.U32 -- 7 multiplier-pipe cycles instead of 4 (SASS +
// tools/ubench3.cu).  So the digit is taken as q'_k = ~lo_k and the row added is (q'_k + 1) * p:
//   * the q'_k * p_j products (j >= 1) are ordinary fused IMAD.WIDEs,
//   * the "+1" parts add the constant sum_i p * 2^(32 i), i.e. c_k = sum of the p_j that fall in column k,
//   * column k's own term lo_k + q'_k * p_0 + p_0 = lo_k + ~lo_k + 1 = 2^32 exactly: the low limb
//     clears and a carry of one moves up, whatever lo_k is (q'_k + 1 ranges over 1..2^32).
// The total added is < 2^256 * p * (1 + 2^-31), so for a*b < p * 2^256 the result is < 2p + 1 (and
// since it is congruent to a*b/2^256 and 2p is not reachable... see cond_sub_r: one conditional subtract
// of p brings any value < 2p to canonical form; the value 2p itself cannot occur because the result
// is < (p^2 + 2^256 p (1 + 2^-31)) / 2^256 < 2p for p < 2^255).
// Inputs as for mont_mul_raw.  REDUCE: out[0..7] (+ top limb returned); else out[0..15] = a * b.
ZKSC_DEV constexpr unsigned long long mod_limb(int j) {
    return j == 0 ? ZKSC_P0 : j == 1 ? ZKSC_P1 : j == 2 ? ZKSC_P2 : j == 3 ? ZKSC_P3 : j == 4 ? ZKSC_P4 : j == 5 ? ZKSC_P5 : j == 6 ? ZKSC_P6 : ZKSC_P7;
}
// sum of the modulus limbs p_j, j >= 1, that the "+1" parts put into column k
ZKSC_DEV constexpr unsigned long long mod_column_constant(int k) {
    unsigned long long c = 0;
    for (int i = (k > 7 ? k - 7 : 0); i <= (k < 8 ? k - 1 : 7); i++) c += mod_limb(k - i);
    return c;
}
template <bool REDUCE>
ZKSC_DEV uint32_t mul_ps(uint32_t* out, const Fr& a, const Fr& b) {
    using namespace ptx;
    // modulus limbs as immediates: IMAD.WIDE.U32 takes an immediate operand, no registers spent
    const uint32_t p[8] = {ZKSC_P0, ZKSC_P1, ZKSC_P2, ZKSC_P3, ZKSC_P4, ZKSC_P5, ZKSC_P6, ZKSC_P7};
    uint32_t q[8];
    uint32_t lo = 0, hi = 0, cn = 0;
#pragma unroll
    for (int k = 0; k < 15; k++) {
#pragma unroll
        for (int i = (k > 7 ? k - 7 : 0); i <= (k < 7 ? k : 7); i++) {
            lo = mad_lo_cc(a.l[i], b.l[k - i], lo);
            hi = madc_hi_cc(a.l[i], b.l[k - i], hi);
            cn = addc(cn, 0u);
        }
        if (REDUCE) {
#pragma unroll
            for (int i = (k > 7 ? k - 7 : 0); i <= (k < 8 ? k - 1 : 7); i++) {
#ifndef ZKSC_NO_P1_TRICK
                if (k - i == 1) {
                    // p_1 = 2^32 - 1:  q * p_1 = (q << 32) - q, five adds on the (idle) ALU pipe instead of
                    // one more IMAD.WIDE on the multiplier pipe that bounds the kernel
                    hi = add_cc(hi, q[i]);
                    cn = addc(cn, 0u);
                    lo = sub_cc(lo, q[i]);
                    hi = subc_cc(hi, 0u);
                    cn = subc(cn, 0u);
                    continue;
                }
#endif
                lo = mad_lo_cc(q[i], p[k - i], lo);
                hi = madc_hi_cc(q[i], p[k - i], hi);
                cn = addc(cn, 0u);
            }
            if (k > 0) {
                const unsigned long long c = mod_column_constant(k);
                lo = add_cc(lo, (uint32_t)c);
                hi = addc_cc(hi, (uint32_t)(c >> 32));
                cn = addc(cn, 0u);
            }
            if (k < 8) {
                q[k] = ~lo;                 // lo + q + 1 == 2^32: limb cleared, carry one
                hi = add_cc(hi, 1u);
                cn = addc(cn, 0u);
            } else {
                out[k - 8] = lo;
            }
        } else {
            out[k] = lo;
        }
        lo = hi; hi = cn; cn = 0;
    }
    if (REDUCE) { out[7] = lo; return hi; }
    out[15] = lo;
    return 0;
}

// ---- operand scanning (row-wise) multiplication with even/odd carry chains ------------------------------
// IMAD.WIDE needs its 64-bit accumulator in an aligned register pair, so partial products whose low limb
// sits at an EVEN position accumulate in E[] and those at an ODD position in O[] (index = absolute limb
// position).  For one multiplier limb x at base position i the products with a0,a2,a4,a6 form one carry
// chain (IMAD.WIDE.U32 / IMAD.WIDE.U32.X) in the array of i's parity and those with a1,a3,a5,a7 a second
// one in the other array: two independent chains per row, and rows overlap, which keeps the multiplier
// pipe busy at 4 warps per scheduler (92% vs 72% for product scanning; profiles/r01_ncu_ubench3_fr_mul.txt).
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
// 512-bit product a*b, any a, b < 2^256
ZKSC_DEV void mul_wide(uint32_t (&T)[16], const Fr& a, const Fr& b) {
    uint32_t E[17], O[17];
#pragma unroll
    for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* A = (i & 1) ? O : E;
        uint32_t* B = (i & 1) ? E : O;
        chain4<false>(A, i, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
        chain4<false>(B, i + 1, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
    }
    T[0] = E[0];
    T[1] = ptx::add_cc(E[1], O[1]);
#pragma unroll
    for (int i = 2; i < 15; i++) T[i] = ptx::addc_cc(E[i], O[i]);
    T[15] = ptx::addc(E[15], O[15]);
}

// ---- one level of Karatsuba on the 8 x 8 limb product --------------------------------------------------------------------
// a = a0 + a1 2^128, b = b0 + b1 2^128:  a b = z0 + (zm - z0 - z2) 2^128 + z2 2^256  with z0 = a0 b0, z2 = a1 b1, zm = (a0 + a1)(b0 + b1):
// three 4 x 4 limb products = 48 IMAD.WIDE instead of 64, paid for with ~70 more ALU instructions (the half sums, their carry bits,
// the two subtractions, the recombination).  Worth it only where the multiplier pipe is the limit and the ALU pipe has room (the
// degree-3 kernels: multiplier pipe 75-80 % busy, ALU pipe 27-32 %; profiles/r02_ncu_c3_round0_and_fold.txt): ZKSC_KARATSUBA.
// two fused lo/hi pairs on X[pos..pos+3], carry into X[pos+4]
ZKSC_DEV void chain2(uint32_t* X, int pos, uint32_t v0, uint32_t v2, uint32_t x) {
    using namespace ptx;
    X[pos + 0] = mad_lo_cc(x, v0, X[pos + 0]);  X[pos + 1] = madc_hi_cc(x, v0, X[pos + 1]);
    X[pos + 2] = madc_lo_cc(x, v2, X[pos + 2]); X[pos + 3] = madc_hi_cc(x, v2, X[pos + 3]);
    X[pos + 4] = addc(X[pos + 4], 0u);
}
// 256-bit product of two 128-bit operands (even / odd IMAD.WIDE chains like mul_wide)
ZKSC_DEV void mul4x4(uint32_t (&T)[8], const uint32_t (&a)[4], const uint32_t (&b)[4]) {
    uint32_t E[9], O[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t* A = (i & 1) ? O : E;
        uint32_t* B = (i & 1) ? E : O;
        chain2(A, i, a[0], a[2], b[i]);
        chain2(B, i + 1, a[1], a[3], b[i]);
    }
    T[0] = E[0];
    T[1] = ptx::add_cc(E[1], O[1]);
#pragma unroll
    for (int i = 2; i < 7; i++) T[i] = ptx::addc_cc(E[i], O[i]);
    T[7] = ptx::addc(E[7], O[7]);
}
// 512-bit product a*b, any a, b < 2^256 (same contract as mul_wide)
ZKSC_DEV void mul_wide_k(uint32_t (&T)[16], const Fr& a, const Fr& b) {
    using namespace ptx;
    uint32_t a0[4], a1[4], b0[4], b1[4], sa[4], sb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { a0[i] = a.l[i]; a1[i] = a.l[4 + i]; b0[i] = b.l[i]; b1[i] = b.l[4 + i]; }
    sa[0] = add_cc(a0[0], a1[0]); sa[1] = addc_cc(a0[1], a1[1]); sa[2] = addc_cc(a0[2], a1[2]); sa[3] = addc_cc(a0[3], a1[3]);
    const uint32_t ca = addc(0u, 0u);                   // a0 + a1 = sa + ca 2^128
    sb[0] = add_cc(b0[0], b1[0]); sb[1] = addc_cc(b0[1], b1[1]); sb[2] = addc_cc(b0[2], b1[2]); sb[3] = addc_cc(b0[3], b1[3]);
    const uint32_t cb = addc(0u, 0u);
    uint32_t z0[8], z2[8], zm[8];
    mul4x4(z0, a0, b0);
    mul4x4(z2, a1, b1);
    mul4x4(zm, sa, sb);
    // middle = zm + (ca sb + cb sa) 2^128 + ca cb 2^256   (< 2^258: ten limbs m[0..9], m[8], m[9] small)
    const uint32_t ma = 0u - ca, mb = 0u - cb;          // all-ones masks
    uint32_t m[10];
#pragma unroll
    for (int i = 0; i < 4; i++) m[i] = zm[i];
    m[4] = add_cc(zm[4], sb[0] & ma); m[5] = addc_cc(zm[5], sb[1] & ma); m[6] = addc_cc(zm[6], sb[2] & ma); m[7] = addc_cc(zm[7], sb[3] & ma);
    m[8] = addc(ca & cb, 0u);
    m[4] = add_cc(m[4], sa[0] & mb); m[5] = addc_cc(m[5], sa[1] & mb); m[6] = addc_cc(m[6], sa[2] & mb); m[7] = addc_cc(m[7], sa[3] & mb);
    m[8] = addc(m[8], 0u);
    // z1 = middle - z0 - z2  (>= 0, < 2^257: nine limbs)
    m[0] = sub_cc(m[0], z0[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) m[i] = subc_cc(m[i], z0[i]);
    m[8] = subc(m[8], 0u);
    m[0] = sub_cc(m[0], z2[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) m[i] = subc_cc(m[i], z2[i]);
    m[8] = subc(m[8], 0u);
    // T = z0 + z1 2^128 + z2 2^256
#pragma unroll
    for (int i = 0; i < 4; i++) T[i] = z0[i];
    T[4] = add_cc(z0[4], m[0]); T[5] = addc_cc(z0[5], m[1]); T[6] = addc_cc(z0[6], m[2]); T[7] = addc_cc(z0[7], m[3]);
    T[8] = addc_cc(z2[0], m[4]); T[9] = addc_cc(z2[1], m[5]); T[10] = addc_cc(z2[2], m[6]); T[11] = addc_cc(z2[3], m[7]);
    T[12] = addc_cc(z2[4], m[8]); T[13] = addc_cc(z2[5], 0u); T[14] = addc_cc(z2[6], 0u); T[15] = addc(z2[7], 0u);
}

// Limb k of C* = sum_{i=0..7} (p - 1 + 2^32) * 2^(32 i): what the "+1" parts of the complement digits add
// (see mont_mul_rows).  Evaluated at compile time.
ZKSC_DEV constexpr uint32_t cstar_limb(int k) {
    unsigned long long carry = 0, limb = 0;
    for (int pos = 0; pos <= k; pos++) {
        unsigned long long s = carry;
        for (int i = 0; i < 8; i++) {
            const int j = pos - i;                       // limb j of (p - 1 + 2^32) lands on position i + j
            if (j < 0 || j > 8) continue;
            // p - 1 + 2^32 = p + 0xffffffff: limbs (p0 - 1 = 0, p1 + 1 -> 0 carry 1, ...) computed on the fly
            unsigned long long c = 0, v = 0;
            for (int l = 0; l <= j; l++) {
                unsigned long long w = (l < 8 ? mod_limb(l) : 0ull) + (l == 0 ? 0xffffffffull : 0ull) + c;
                v = w & 0xffffffffull;
                c = w >> 32;
            }
            s += v;
        }
        limb = s & 0xffffffffull;
        carry = s >> 32;
    }
    return (uint32_t)limb;
}

template <int K>
struct CStar {
    static constexpr uint32_t v = cstar_limb(K);   // forces compile-time evaluation
};

// Montgomery multiplication (CIOS, even/odd chains, complement digits).
// Row i adds a*b_i, then reads limb i of the running total, t = E[i] + O[i] + k (k = carry out of limb
// i-1), and adds (m' + 1) * p * 2^(32 i) with m' = ~t:
//   * m' is a COMPLEMENT, not the textbook negation -t: ptxas folds a negation into operand modifiers and
//     then splits every digit product into IMAD + IMAD.HI.U32 (7 multiplier-pipe cycles instead of ~4);
//   * limb i becomes t + m' * p_0 = 2^32 - 1, and its "+1" would turn that into 0 with a carry of one, so
//     the +1 is booked one limb higher instead; all the "+1" parts together are the constant C* that E[]
//     starts from (no run-time cost).  Limbs 0..7 end as 0xffffffff instead of 0 and are discarded;
//   * m' * p_0 = m' needs no multiplication: an add at the head of the even chain.
// 64 + 56 IMAD.WIDE in all.  Inputs: a, b < 2^256 with a*b < p * 2^256 for a result < 2p.
// Returns limbs 8..16 of the total: res[0..7] and the top limb.
ZKSC_DEV void mont_mul_rows(uint32_t (&res)[8], uint32_t& top, const Fr& a, const Fr& b) {
    using namespace ptx;
    uint32_t E[18], O[18];
#pragma unroll
    for (int i = 0; i < 18; i++) O[i] = 0;
    E[0] = CStar<0>::v; E[1] = CStar<1>::v; E[2] = CStar<2>::v; E[3] = CStar<3>::v; E[4] = CStar<4>::v; E[5] = CStar<5>::v;
    E[6] = CStar<6>::v; E[7] = CStar<7>::v; E[8] = CStar<8>::v; E[9] = CStar<9>::v; E[10] = CStar<10>::v; E[11] = CStar<11>::v;
    E[12] = CStar<12>::v; E[13] = CStar<13>::v; E[14] = CStar<14>::v; E[15] = CStar<15>::v; E[16] = CStar<16>::v; E[17] = 0u;
    const uint32_t p1 = ZKSC_P1, p2 = ZKSC_P2, p3 = ZKSC_P3, p4 = ZKSC_P4, p5 = ZKSC_P5, p6 = ZKSC_P6, p7 = ZKSC_P7;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* A = (i & 1) ? O : E;
        uint32_t* B = (i & 1) ? E : O;
        chain4<false>(A, i, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
        chain4<false>(B, i + 1, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
        uint32_t m;
        if (i == 0) {
            m = ~E[0];
            A[0] = add_cc(A[0], m);
        } else {
            (void)add_cc(E[i - 1], O[i - 1]);          // limb i-1 is 2^32 - 1 (+ 2^32 k): CF = k
            m = ~addc(E[i], O[i]);
            A[i] = addc_cc(A[i], m);                   // k enters here
        }
        A[i + 1] = addc_cc(A[i + 1], 0u);
        A[i + 2] = madc_lo_cc(m, p2, A[i + 2]); A[i + 3] = madc_hi_cc(m, p2, A[i + 3]);
        A[i + 4] = madc_lo_cc(m, p4, A[i + 4]); A[i + 5] = madc_hi_cc(m, p4, A[i + 5]);
        A[i + 6] = madc_lo_cc(m, p6, A[i + 6]); A[i + 7] = madc_hi_cc(m, p6, A[i + 7]);
        A[i + 8] = addc(A[i + 8], 0u);
        chain4<false>(B, i + 1, p1, p3, p5, p7, m);
    }
    (void)add_cc(O[7], E[7]);                          // k_7
#pragma unroll
    for (int i = 0; i < 8; i++) res[i] = addc_cc(E[8 + i], O[8 + i]);
    top = addc(E[16], O[16]);
}

// The same Montgomery product with the multiplication done first (mul_wide_k) and the eight digit rows afterwards: row i of
// mont_mul_rows reads limb i of the running total, and the product rows j > i only touch limbs >= j, so the digits -- and the
// result -- are the ones mont_mul_rows finds.  48 + 56 IMAD.WIDE instead of 64 + 56.
ZKSC_DEV void mont_mul_rows_k(uint32_t (&res)[8], uint32_t& top, const Fr& a, const Fr& b) {
    using namespace ptx;
    uint32_t T[16];
    mul_wide_k(T, a, b);
    uint32_t E[18], O[18];
#pragma unroll
    for (int i = 0; i < 18; i++) O[i] = 0;
    E[0] = add_cc(T[0], CStar<0>::v); E[1] = addc_cc(T[1], CStar<1>::v); E[2] = addc_cc(T[2], CStar<2>::v); E[3] = addc_cc(T[3], CStar<3>::v);
    E[4] = addc_cc(T[4], CStar<4>::v); E[5] = addc_cc(T[5], CStar<5>::v); E[6] = addc_cc(T[6], CStar<6>::v); E[7] = addc_cc(T[7], CStar<7>::v);
    E[8] = addc_cc(T[8], CStar<8>::v); E[9] = addc_cc(T[9], CStar<9>::v); E[10] = addc_cc(T[10], CStar<10>::v); E[11] = addc_cc(T[11], CStar<11>::v);
    E[12] = addc_cc(T[12], CStar<12>::v); E[13] = addc_cc(T[13], CStar<13>::v); E[14] = addc_cc(T[14], CStar<14>::v); E[15] = addc_cc(T[15], CStar<15>::v);
    E[16] = addc(0u, CStar<16>::v); E[17] = 0u;
    const uint32_t p1 = ZKSC_P1, p2 = ZKSC_P2, p3 = ZKSC_P3, p4 = ZKSC_P4, p5 = ZKSC_P5, p6 = ZKSC_P6, p7 = ZKSC_P7;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* A = (i & 1) ? O : E;
        uint32_t* B = (i & 1) ? E : O;
        uint32_t m;
        if (i == 0) {
            m = ~E[0];
            A[0] = add_cc(A[0], m);
        } else {
            (void)add_cc(E[i - 1], O[i - 1]);          // limb i-1 is 2^32 - 1 (+ 2^32 k): CF = k
            m = ~addc(E[i], O[i]);
            A[i] = addc_cc(A[i], m);                   // k enters here
        }
        A[i + 1] = addc_cc(A[i + 1], 0u);
        A[i + 2] = madc_lo_cc(m, p2, A[i + 2]); A[i + 3] = madc_hi_cc(m, p2, A[i + 3]);
        A[i + 4] = madc_lo_cc(m, p4, A[i + 4]); A[i + 5] = madc_hi_cc(m, p4, A[i + 5]);
        A[i + 6] = madc_lo_cc(m, p6, A[i + 6]); A[i + 7] = madc_hi_cc(m, p6, A[i + 7]);
        A[i + 8] = addc(A[i + 8], 0u);
        chain4<false>(B, i + 1, p1, p3, p5, p7, m);
    }
    (void)add_cc(O[7], E[7]);                          // k_7
#pragma unroll
    for (int i = 0; i < 8; i++) res[i] = addc_cc(E[8 + i], O[8 + i]);
    top = addc(E[16], O[16]);
}

// Montgomery product, canonical result.  a, b < r.
ZKSC_DEV Fr fr_mul(const Fr& a, const Fr& b) {
    Fr o;
    uint32_t top;
#ifdef ZKSC_MUL_PRODUCT_SCANNING
    top = mul_ps<true>(o.l, a, b);
#else
    mont_mul_rows(o.l, top, a, b);   // < 2r < 2^256: top == 0
#endif
    (void)top;
    cond_sub_r(o.l);
    return o;
}

// Montgomery product WITHOUT the final conditional subtraction: a * b < r * 2^256  ->  a value < 2r congruent to a b / R.
// For a product that goes straight into mul_wide (which takes any 256-bit operand): 17 ALU instructions saved.
ZKSC_DEV Fr fr_mul_lazy(const Fr& a, const Fr& b) {
    Fr o;
    uint32_t top;
    mont_mul_rows(o.l, top, a, b);   // < 2r < 2^256: top == 0
    (void)top;
    return o;
}

// the product-first (Karatsuba) forms of fr_mul / fr_mul_lazy
ZKSC_DEV Fr fr_mul_k(const Fr& a, const Fr& b) {
    Fr o;
    uint32_t top;
    mont_mul_rows_k(o.l, top, a, b);
    (void)top;
    cond_sub_r(o.l);
    return o;
}
ZKSC_DEV Fr fr_mul_lazy_k(const Fr& a, const Fr& b) {
    Fr o;
    uint32_t top;
    mont_mul_rows_k(o.l, top, a, b);
    (void)top;
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
            unsigned long long w = (unsigned long long)m * mod_limb(j) + T[i + j] + c;
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

#ifndef ZKSC_HOST_EMU
// The same, called by a WHOLE warp with the sum in lane 0: the two Montgomery products run side by side in lanes 0 and 1 (one
// instruction stream, two operand sets) instead of one after the other -- a lone warp finishing a round's sum is pure dependency
// latency (~4 cycles per instruction of the carry chains), so this halves the ~2 us the serial form cost every round
// (profiles/r02_trace_resident_c2_v2.txt).  The result is meaningful in lane 0.
ZKSC_DEV Fr acc17_reduce_warp(const Acc<17>& a) {
    const int lane = threadIdx.x & 31;
    Fr lo, mid, x = fr_zero(), y = fr_zero();
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.l[i] = a.l[i]; mid.l[i] = a.l[8 + i]; }
    const uint32_t h = __shfl_sync(0xffffffffu, a.l[16], 0);
    if (lane == 0) { x = fr_canon(lo); y.l[0] = 1u; }
    if (lane == 1) { x.l[0] = h; y = fr_r2(); }
    Fr r = fr_mul(x, y);
    Fr r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = __shfl_sync(0xffffffffu, r.l[i], 1);
    return fr_add(fr_add(r, fr_canon(mid)), r2);
}
#endif

// fold: a + r * (b - a)   (== r*b + (1-r)*a, polynomial/src/multilinear/evaluation_form.rs:133)
ZKSC_DEV Fr fr_fold(const Fr& a, const Fr& b, const Fr& r) {
    return fr_add(a, fr_mul(r, fr_sub(b, a)));
}

// ---- multiplication by a FIXED element through a precomputed shift table --------------------------
// A whole round folds every table entry with the SAME challenge r, so the host precomputes (once per
// round, 8 Montgomery multiplications)
//     w[i] = r * 2^(32 (i + 2)) mod p,   i = 0..7      (canonical residues; r = the plain challenge)
// and the device evaluates  r * d  for d = sum d_i 2^(32 i)  as  T = sum_i d_i * w[i]  -- an 8 x 8 limb
// product whose rows all land on limb 0 (the shifts live in the table), T < 2^35 p < 2^290 -- followed by
// TWO Montgomery digit steps that divide out the extra 2^64:  V = (T + M p) / 2^64 < 2^226 + p(1 + 2^-32)
// < 2p,  V == r * d (mod p).  With d in Montgomery form, V is the Montgomery form of r * (d/R): exactly
// what mont_mul(r~, d) returns, at 64 + 12 IMAD.WIDE instead of 64 + 56.
// Digit steps.  p = 1 - 2^32 + 2^64 K  with  K = (p >> 64) + 1  (six limbs), so for a digit m' = ~t (t = the
// low limb, complement instead of negation for the reason given at mont_mul_rows):
//     (T + (m' + 1) p) / 2^32 = (T >> 32) + (t + 1) + 2^32 (m' K + K - 1)
// -- every term non-negative, six products per digit.  Both digits follow from limbs 0 and 1 of T alone
// (u = t_1 + t_0 + 1 is the low limb after step one), so the 12 reduction products do not wait for
// each other.  K - 1 = p >> 64 of both steps is the constant E[] starts from.
struct FoldTab {
    uint32_t w[8][8];
};
ZKSC_DEV constexpr uint32_t foldk_limb(int j) { return (uint32_t)(mod_limb(j + 2) + (j == 0 ? 1ull : 0ull)); }   // K, limbs 0..5
// limb k (absolute position, k >= 2) of ((p >> 64) * (1 + 2^32)) << 64
ZKSC_DEV constexpr uint32_t foldq_limb(int k) {
    unsigned long long carry = 0, limb = 0;
    for (int pos = 2; pos <= k; pos++) {
        unsigned long long s = carry;
        const int j0 = pos - 2, j1 = pos - 3;            // Q limb j0 from the first copy, j1 from the shifted copy
        if (j0 >= 0 && j0 < 6) s += mod_limb(j0 + 2);
        if (j1 >= 0 && j1 < 6) s += mod_limb(j1 + 2);
        limb = s & 0xffffffffull;
        carry = s >> 32;
    }
    return (uint32_t)limb;
}
template <int K>
struct FoldQ {
    static constexpr uint32_t v = foldq_limb(K);
};
// three fused lo/hi pairs on X[pos..pos+5], carry into X[pos+6]
ZKSC_DEV void chain3(uint32_t* X, int pos, uint32_t v0, uint32_t v2, uint32_t v4, uint32_t x) {
    using namespace ptx;
    X[pos + 0] = mad_lo_cc(x, v0, X[pos + 0]);  X[pos + 1] = madc_hi_cc(x, v0, X[pos + 1]);
    X[pos + 2] = madc_lo_cc(x, v2, X[pos + 2]); X[pos + 3] = madc_hi_cc(x, v2, X[pos + 3]);
    X[pos + 4] = madc_lo_cc(x, v4, X[pos + 4]); X[pos + 5] = madc_hi_cc(x, v4, X[pos + 5]);
    X[pos + 6] = addc(X[pos + 6], 0u);
}
// The same table for kernels that receive the challenge AFTER they were launched (the resident rounds kernel): it lives in
// shared memory, rows permuted to (w0, w2, w4, w6 | w1, w3, w5, w7) so that the four multipliers of each carry chain arrive
// with one LDS.128.
struct FoldTabS {
    uint32_t w[64];
};
#ifndef ZKSC_HOST_EMU
__host__ __device__
#endif
constexpr int foldtabs_index(int i, int j) { return i * 8 + (j & 1) * 4 + (j >> 1); }   // where w[i][j] of a FoldTab sits
ZKSC_DEV void fold_row(const FoldTab& W, int i, uint32_t (&e)[4], uint32_t (&o)[4]) {
    e[0] = W.w[i][0]; e[1] = W.w[i][2]; e[2] = W.w[i][4]; e[3] = W.w[i][6];
    o[0] = W.w[i][1]; o[1] = W.w[i][3]; o[2] = W.w[i][5]; o[3] = W.w[i][7];
}
#ifndef ZKSC_HOST_EMU
ZKSC_DEV void fold_row(const FoldTabS& W, int i, uint32_t (&e)[4], uint32_t (&o)[4]) {
    const uint4 a = *reinterpret_cast<const uint4*>(&W.w[8 * i]), b = *reinterpret_cast<const uint4*>(&W.w[8 * i + 4]);
    e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w;
    o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w;
}
#else
ZKSC_DEV void fold_row(const FoldTabS& W, int i, uint32_t (&e)[4], uint32_t (&o)[4]) {
    for (int k = 0; k < 4; k++) { e[k] = W.w[8 * i + k]; o[k] = W.w[8 * i + 4 + k]; }
}
#endif
// res (8 limbs, < 2p) == r * d (mod p) for ANY d < 2^256, W the shift table of r.
template <class TAB>
ZKSC_DEV void mul_fixed_rows(uint32_t (&res)[8], const uint32_t (&d)[8], const TAB& W) {
    using namespace ptx;
    uint32_t E[11], O[11];
#pragma unroll
    for (int i = 0; i < 11; i++) O[i] = 0;
    E[0] = 0; E[1] = 0; E[2] = FoldQ<2>::v; E[3] = FoldQ<3>::v; E[4] = FoldQ<4>::v; E[5] = FoldQ<5>::v;
    E[6] = FoldQ<6>::v; E[7] = FoldQ<7>::v; E[8] = FoldQ<8>::v; E[9] = FoldQ<9>::v; E[10] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t we[4], wo[4];
        fold_row(W, i, we, wo);
        chain4<false>(E, 0, we[0], we[1], we[2], we[3], d[i]);
        chain4<false>(O, 1, wo[0], wo[1], wo[2], wo[3], d[i]);
    }
    // limbs 0 and 1 of T, the two digits, and what the low 64 bits leave behind at limb 2
    const uint32_t t0 = E[0];
    const uint32_t t1 = add_cc(E[1], O[1]);
    uint32_t cs = addc(0u, 0u);                       // c1: carry of E + O out of limb 1
    uint32_t u = add_cc(t1, t0);
    cs = addc(cs, 0u);
    u = add_cc(u, 1u);
    cs = addc(cs, 0u);                                // + c0: carry of t1 + t0 + 1
    uint32_t small_lo = add_cc(u, 1u);                // small = (u + 1) + c0 + c1, at limb 2
    uint32_t small_hi = addc(0u, 0u);
    small_lo = add_cc(small_lo, cs);
    small_hi = addc(small_hi, 0u);
    const uint32_t m0 = ~t0, m1 = ~u;
    const uint32_t k0 = foldk_limb(0), k1 = foldk_limb(1), k2 = foldk_limb(2), k3 = foldk_limb(3), k4 = foldk_limb(4), k5 = foldk_limb(5);
    chain3(E, 2, k0, k2, k4, m0);
    chain3(O, 3, k1, k3, k5, m0);
    chain3(O, 3, k0, k2, k4, m1);
    chain3(E, 4, k1, k3, k5, m1);
    res[0] = add_cc(E[2], O[2]);
#pragma unroll
    for (int i = 1; i < 7; i++) res[i] = addc_cc(E[2 + i], O[2 + i]);
    res[7] = addc(E[9], O[9]);
    res[0] = add_cc(res[0], small_lo);
    res[1] = addc_cc(res[1], small_hi);
#pragma unroll
    for (int i = 2; i < 7; i++) res[i] = addc_cc(res[i], 0u);
    res[7] = addc(res[7], 0u);
}
// a < p, v < p + 2^227 (what mul_fixed_rows returns: v < p + 2^226 + p 2^-32)  ->  canonical (a + v) mod p.
// a + v < 2p + 2^227 < 2^256, so there is no carry out; ONE conditional subtraction leaves s < p + 2^227, and s >= p is then
// possible only for s in [p, p + 2^227) -- never for practical purposes, but exactness is unconditional: s >= p implies that
// its top limb is >= p's top limb, and that (almost never taken) branch finishes the job out of line.  17 ALU instructions
// fewer per fold than cond_sub_r(v) followed by fr_add(a, v).
ZKSC_DEV Fr fr_add_semi(const Fr& a, const Fr& v) {
    using namespace ptx;
    Fr s;
    s.l[0] = add_cc(a.l[0], v.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) s.l[i] = addc_cc(a.l[i], v.l[i]);
    s.l[7] = addc(a.l[7], v.l[7]);
    cond_sub_r(s.l);
    if (__builtin_expect(s.l[7] >= ZKSC_P7, 0)) s = cond_sub_r_rare(s);
    return s;
}
// fold through the table: a + r * (b - a), canonical.  a, b < p.
// SEMI: finish with fr_add_semi (one conditional subtraction + a never-taken branch) instead of cond_sub_r + fr_add.  17 ALU
// instructions fewer, but measured faster only where ptxas has registers to spare: round_kernel<2, FOLD, SKIP1, 1> 332.5 -> 316.7 us
// (c2 round 1), while the d = 3 and the 64-proof instantiations lose 1.5-2.7 % to it (profiles/r01_variants_v5.txt).
template <bool SEMI = false, class TAB = FoldTab>
ZKSC_DEV Fr fr_fold_tab(const Fr& a, const Fr& b, const TAB& W) {
    using namespace ptx;
    uint32_t d[8];                                     // b - a + p  in (0, 2p): no conditional
    d[0] = sub_cc(b.l[0], a.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) d[i] = subc_cc(b.l[i], a.l[i]);
    d[7] = subc(b.l[7], a.l[7]);
    d[0] = add_cc(d[0], ZKSC_P0); d[1] = addc_cc(d[1], ZKSC_P1); d[2] = addc_cc(d[2], ZKSC_P2); d[3] = addc_cc(d[3], ZKSC_P3);
    d[4] = addc_cc(d[4], ZKSC_P4); d[5] = addc_cc(d[5], ZKSC_P5); d[6] = addc_cc(d[6], ZKSC_P6); d[7] = addc(d[7], ZKSC_P7);
    Fr v;
    mul_fixed_rows(v.l, d, W);
    if constexpr (SEMI) {
        return fr_add_semi(a, v);
    } else {
        cond_sub_r(v.l);
        return fr_add(a, v);
    }
}

}  // namespace zksc
