// TEST INFRASTRUCTURE ONLY.  Compiles zk_cryptography_b200/csrc/fr.cuh for the host with the PTX
// extended-precision instructions emulated (ZKSC_HOST_EMU), so the exact limb/carry-chain algorithms
// that run on the GPU can be checked against Python big integers on a CPU-only box.  Never linked
// into the product library.
#define ZKSC_HOST_EMU 1
#include "../../zk_cryptography_b200/csrc/fr.cuh"
#include <cstring>
using namespace zksc;
extern "C" {
void emu_mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* out16) {
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    uint32_t T[16]; mul_wide(T, x, y); memcpy(out16, T, 64);
}
void emu_redc(const uint32_t* t16, uint32_t* out9) {
    uint32_t T[16], top; memcpy(T, t16, 64); redc_rows(T, top); memcpy(out9, T + 8, 32); out9[8] = top;
}
#define BIN(name, fn) void name(const uint32_t* a, const uint32_t* b, uint32_t* o) { Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); Fr z = fn(x, y); memcpy(o, z.l, 32); }
BIN(emu_fr_mul, fr_mul)
BIN(emu_fr_add, fr_add)
BIN(emu_fr_sub, fr_sub)
void emu_fr_canon(const uint32_t* a, uint32_t* o) { Fr x; memcpy(x.l, a, 32); Fr z = fr_canon(x); memcpy(o, z.l, 32); }
void emu_fr_fold(const uint32_t* a, const uint32_t* b, const uint32_t* r, uint32_t* o) {
    Fr x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32); memcpy(z.l, r, 32); Fr w = fr_fold(x, y, z); memcpy(o, w.l, 32);
}
void emu_acc9_reduce(const uint32_t* a9, uint32_t* o) { Acc<9> a; memcpy(a.l, a9, 36); Fr z = acc9_reduce(a); memcpy(o, z.l, 32); }
void emu_acc17_reduce(const uint32_t* a17, uint32_t* o) { Acc<17> a; memcpy(a.l, a17, 68); Fr z = acc17_reduce(a); memcpy(o, z.l, 32); }
void emu_acc17_add(uint32_t* a17, const uint32_t* t16) { Acc<17> a; memcpy(a.l, a17, 68); uint32_t T[16]; memcpy(T, t16, 64); acc_add<17, 16>(a, T); memcpy(a17, a.l, 68); }
void emu_consts(uint32_t* one, uint32_t* r2) { Fr a = fr_one(), b = fr_r2(); memcpy(one, a.l, 32); memcpy(r2, b.l, 32); }
}
extern "C" void emu_mul_ps_wide(const uint32_t* a, const uint32_t* b, uint32_t* out16) {
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    uint32_t T[16]; (void)mul_ps<false>(T, x, y); memcpy(out16, T, 64);
}
extern "C" void emu_mul_ps_mont(const uint32_t* a, const uint32_t* b, uint32_t* out9) {
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    uint32_t res[8]; uint32_t top = mul_ps<true>(res, x, y); memcpy(out9, res, 32); out9[8] = top;
}
extern "C" void emu_mont_mul_rows(const uint32_t* a, const uint32_t* b, uint32_t* out9) {
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    uint32_t res[8], top; mont_mul_rows(res, top, x, y); memcpy(out9, res, 32); out9[8] = top;
}
extern "C" void emu_cstar(uint32_t* out17) { for (int i = 0; i < 17; i++) out17[i] = cstar_limb(i); }
extern "C" void emu_mul_fixed(const uint32_t* d, const uint32_t* w64, uint32_t* out8) {
    FoldTab W; memcpy(W.w, w64, 256);
    uint32_t dd[8], res[8]; memcpy(dd, d, 32);
    mul_fixed_rows(res, dd, W); memcpy(out8, res, 32);
}
extern "C" void emu_fr_fold_tab(const uint32_t* a, const uint32_t* b, const uint32_t* w64, uint32_t* o) {
    FoldTab W; memcpy(W.w, w64, 256);
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); Fr z = fr_fold_tab(x, y, W); memcpy(o, z.l, 32);
}
extern "C" {
BIN(emu_fr_add_semi, fr_add_semi)
BIN(emu_fr_add_lazy, fr_add_lazy)
}
extern "C" void emu_fr_fold_tab_semi(const uint32_t* a, const uint32_t* b, const uint32_t* w64, uint32_t* o) {
    FoldTab W; memcpy(W.w, w64, 256);
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); Fr z = fr_fold_tab<true>(x, y, W); memcpy(o, z.l, 32);
}
// the resident kernel's form: the same table in shared-memory order (FoldTabS, rows permuted for LDS.128)
extern "C" void emu_fr_fold_tabs_semi(const uint32_t* a, const uint32_t* b, const uint32_t* w64, uint32_t* o) {
    FoldTabS W;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) W.w[foldtabs_index(i, j)] = w64[8 * i + j];
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); Fr z = fr_fold_tab<true>(x, y, W); memcpy(o, z.l, 32);
}
extern "C" {
BIN(emu_fr_mul_lazy, fr_mul_lazy)
}
extern "C" {
BIN(emu_fr_sub_lazy, fr_sub_lazy)
}

// one level of Karatsuba (mul_wide_k) and the product-first Montgomery multiplication built on it (mont_mul_rows_k)
extern "C" void emu_mul_wide_k(const uint32_t* a, const uint32_t* b, uint32_t* out16) {
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    uint32_t T[16]; mul_wide_k(T, x, y); memcpy(out16, T, 64);
}
extern "C" void emu_mont_mul_rows_k(const uint32_t* a, const uint32_t* b, uint32_t* out9) {
    Fr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    uint32_t res[8], top; mont_mul_rows_k(res, top, x, y); memcpy(out9, res, 32); out9[8] = top;
}
