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

// Montgomery product, canonical result.  a, b < r.
ZKSC_DEV Fr fr_mul(const Fr& a, const Fr& b) {
    Fr o;
    (void)mul_ps<true>(o.l, a, b);   // < 2r < 2^256: top == 0
    cond_sub_r(o.l);
    return o;
}
// 512-bit product a*b, any a, b < 2^256
ZKSC_DEV void mul_wide(uint32_t (&T)[16], const Fr& a, const Fr& b) { (void)mul_ps<false>(T, a, b); }

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

// fold: a + r * (b - a)   (== r*b + (1-r)*a, polynomial/src/multilinear/evaluation_form.rs:133)
ZKSC_DEV Fr fr_fold(const Fr& a, const Fr& b, const Fr& r) {
    return fr_add(a, fr_mul(r, fr_sub(b, a)));
}

}  // namespace zksc
