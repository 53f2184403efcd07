"""The limb / carry-chain algorithms of zk_cryptography_b200/csrc/fr.cuh, compiled for the host with the
PTX extended-precision instructions emulated (tests/emu/fr_emu.cpp), against Python big integers.  This
checks on a CPU-only box the exact algorithms the GPU executes (the -m gpu tests check the real thing)."""
import ctypes
import os
import random
import subprocess

import pytest

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
RR = 1 << 256
RINV = pow(RR, -1, R)
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(HERE, "emu", "libfr_emu.so")
    src = os.path.join(HERE, "emu", "fr_emu.cpp")
    hdr = os.path.join(HERE, "..", "zk_cryptography_b200", "csrc", "fr.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)


def arr(x, n):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


EDGE = [0, 1, R - 1, R - 2, 2**256 - 1, 2**255, (1 << 32) - 1, R + 1, 2**256 - 2**32, 0xFFFFFFFF << 224, R // 2, 2 * R - 1]


def rnd(rng, lim):
    if rng.random() < 0.25:
        return rng.choice(EDGE) % lim
    return rng.randrange(lim)


def test_constants(emu):
    one, r2 = (ctypes.c_uint32 * 8)(), (ctypes.c_uint32 * 8)()
    emu.emu_consts(one, r2)
    assert val(one) == RR % R and val(r2) == RR * RR % R


def test_complement_digit_constant(emu):
    c = (ctypes.c_uint32 * 17)()
    emu.emu_cstar(c)
    assert val(c) == sum((R - 1 + 2**32) << (32 * i) for i in range(8))


def test_limb_algorithms(emu):
    rng = random.Random(11)
    o8, o9, o16 = (ctypes.c_uint32 * 8)(), (ctypes.c_uint32 * 9)(), (ctypes.c_uint32 * 16)()
    for it in range(6000):
        a, b = rnd(rng, 2**256), rnd(rng, 2**256)
        emu.emu_mul_wide(arr(a, 8), arr(b, 8), o16)
        assert val(o16) == a * b
        emu.emu_mul_ps_wide(arr(a, 8), arr(b, 8), o16)
        assert val(o16) == a * b
        emu.emu_mul_ps_mont(arr(a, 8), arr(b, 8), o9)
        v = val(o9)
        assert v % R == a * b * RINV % R and v < 2**256 + R
        emu.emu_mont_mul_rows(arr(a, 8), arr(b, 8), o9)
        v = val(o9)
        assert v % R == a * b * RINV % R and v < 2**256 + R
        T = rnd(rng, 2**512)
        emu.emu_redc(arr(T, 16), o9)
        v = val(o9)
        assert v % R == T * RINV % R and v < 2**256 + R
        a %= R
        b %= R
        emu.emu_fr_mul(arr(a, 8), arr(b, 8), o8); assert val(o8) == a * b * RINV % R
        emu.emu_fr_add(arr(a, 8), arr(b, 8), o8); assert val(o8) == (a + b) % R
        emu.emu_fr_sub(arr(a, 8), arr(b, 8), o8); assert val(o8) == (a - b) % R
        x = rnd(rng, 2**256)
        emu.emu_fr_canon(arr(x, 8), o8); assert val(o8) == x % R
        c = rnd(rng, R)
        emu.emu_fr_fold(arr(a, 8), arr(b, 8), arr(c, 8), o8); assert val(o8) == (a + c * (b - a) * RINV) % R
        A = rnd(rng, 2**288)
        emu.emu_acc9_reduce(arr(A, 9), o8); assert val(o8) == A % R
        A = rnd(rng, 2**544)
        emu.emu_acc17_reduce(arr(A, 17), o8); assert val(o8) == A * RINV % R


def test_lazy_accumulation_of_many_products(emu):
    """sum of unreduced 512-bit products in a 17-limb accumulator, one reduction at the end"""
    rng = random.Random(12)
    acc = (ctypes.c_uint32 * 17)()
    total = 0
    o16, o8 = (ctypes.c_uint32 * 16)(), (ctypes.c_uint32 * 8)()
    for _ in range(500):
        a, b = rnd(rng, 2**256), rnd(rng, 2**256)   # operands need not be canonical
        emu.emu_mul_wide(arr(a, 8), arr(b, 8), o16)
        emu.emu_acc17_add(acc, o16)
        total += a * b
    assert val(acc) == total
    emu.emu_acc17_reduce(acc, o8)
    assert val(o8) == total * RINV % R


def fold_table(r_plain):
    """w[i] = r * 2^(32 (i + 2)) mod p as 8 x 8 little-endian 32-bit limbs (fr.cuh FoldTab)"""
    flat = []
    for i in range(8):
        w = r_plain * (1 << (32 * (i + 2))) % R
        flat += [(w >> (32 * j)) & 0xFFFFFFFF for j in range(8)]
    return (ctypes.c_uint32 * 64)(*flat)


def test_fixed_multiplicand_fold(emu):
    """mul_fixed_rows / fr_fold_tab: r * d through the per-round shift table, two Montgomery digits"""
    rng = random.Random(13)
    o8 = (ctypes.c_uint32 * 8)()
    for it in range(4000):
        r = rnd(rng, R)
        W = fold_table(r)
        d = rnd(rng, 2**256)
        emu.emu_mul_fixed(arr(d, 8), W, o8)
        v = val(o8)
        assert v % R == r * d % R and v < 2 * R
        a, b = rnd(rng, R), rnd(rng, R)
        emu.emu_fr_fold_tab(arr(a, 8), arr(b, 8), W, o8)
        assert val(o8) == (a + r * (b - a)) % R
        emu.emu_fr_fold_tab_semi(arr(a, 8), arr(b, 8), W, o8)        # fr_fold_tab<true>: single conditional subtraction
        assert val(o8) == (a + r * (b - a)) % R
        # the bound fr_add_semi relies on: mul_fixed_rows stays under r + 2^227
        assert v < R + 2**227
    # folds of extreme operands through both tails
    for r in (0, 1, R - 1, R // 2):
        W = fold_table(r)
        for a in (0, 1, R - 1):
            for b in (0, 1, R - 1, R - 2):
                for fn in (emu.emu_fr_fold_tab, emu.emu_fr_fold_tab_semi, emu.emu_fr_fold_tabs_semi):
                    fn(arr(a, 8), arr(b, 8), W, o8)
                    assert val(o8) == (a + r * (b - a)) % R
    # extreme table / operand limbs
    for r in (0, 1, R - 1):
        W = fold_table(r)
        for d in (0, 1, 2**256 - 1, 2**32 - 1, 2**64 - 1, R, R - 1):
            emu.emu_mul_fixed(arr(d, 8), W, o8)
            v = val(o8)
            assert v % R == r * d % R and v < 2 * R


def test_semi_reduced_add_and_lazy_add(emu):
    """fr_add_semi (the tail of fr_fold_tab): a < r, v < r + 2^227 -> canonical (a + v) mod r, including the out-of-line branch for
    v >= r that random folds practically never reach; fr_add_lazy: plain 256-bit sum, no reduction."""
    rng = random.Random(77)
    o8 = (ctypes.c_uint32 * 8)()
    edge_a = [0, 1, R - 1, R - 2, R // 2]
    edge_v = [0, 1, R - 1, R, R + 1, R + 2**226, R + 2**227 - 1, 2**227]
    cases = [(a, v) for a in edge_a for v in edge_v]
    cases += [(rnd(rng, R), rnd(rng, R + 2**227)) for _ in range(300)]
    cases += [(R - 1 - rnd(rng, 2**40), R + rnd(rng, 2**227)) for _ in range(300)]      # a + v >= 2r: needs the second subtraction
    for a, v in cases:
        emu.emu_fr_add_semi(arr(a, 8), arr(v, 8), o8)
        assert val(o8) == (a + v) % R
    for _ in range(200):
        a, b = rnd(rng, R), rnd(rng, R)
        emu.emu_fr_add_lazy(arr(a, 8), arr(b, 8), o8)
        assert val(o8) == a + b


def test_lazy_montgomery_product(emu):
    """fr_mul_lazy: no final conditional subtraction; for a < r, b < 2r (a lazy sum) the result is < 2r and congruent to a b / R"""
    rng = random.Random(91)
    o8 = (ctypes.c_uint32 * 8)()
    cases = [(rnd(rng, R), rnd(rng, 2 * R)) for _ in range(2000)] + [(R - 1, 2 * R - 1), (R - 1, 2 * R - 2), (0, 2 * R - 1), (1, 1), (R - 1, R - 1)]
    for a, b in cases:
        emu.emu_fr_mul_lazy(arr(a, 8), arr(b, 8), o8)
        v = val(o8)
        assert v < 2 * R and v % R == a * b * RINV % R


def test_lazy_difference(emu):
    """fr_sub_lazy: a - b + r as a plain 256-bit value in (0, 2r), no borrow test"""
    rng = random.Random(92)
    o8 = (ctypes.c_uint32 * 8)()
    for a, b in [(rnd(rng, R), rnd(rng, R)) for _ in range(500)] + [(0, R - 1), (R - 1, 0), (0, 0), (R - 1, R - 1), (1, 2)]:
        emu.emu_fr_sub_lazy(arr(a, 8), arr(b, 8), o8)
        assert val(o8) == a - b + R


def test_karatsuba_product_and_product_first_montgomery(emu):
    """mul_wide_k (one level of Karatsuba: three 4 x 4 limb products, the half sums' carry bits, two subtractions) is the full
    512-bit product for ANY 256-bit operands, and mont_mul_rows_k (that product first, the eight Montgomery digit rows afterwards)
    returns exactly what mont_mul_rows returns -- same digits, same limbs"""
    rng = random.Random(77)
    o9, p9, o16 = (ctypes.c_uint32 * 9)(), (ctypes.c_uint32 * 9)(), (ctypes.c_uint32 * 16)()
    halves = [0, 1, 2**128 - 1, 2**127, 2**64, 2**128 - 2**32, (1 << 96) - 1]
    cases = [(lo + (hi << 128), lo2 + (hi2 << 128)) for lo in halves for hi in halves[:4] for lo2 in halves[:4] for hi2 in halves]   # carries out of both half sums
    cases += [(rnd(rng, 2**256), rnd(rng, 2**256)) for _ in range(6000)]
    for a, b in cases:
        emu.emu_mul_wide_k(arr(a, 8), arr(b, 8), o16)
        assert val(o16) == a * b, (hex(a), hex(b))
        emu.emu_mont_mul_rows_k(arr(a, 8), arr(b, 8), o9)
        emu.emu_mont_mul_rows(arr(a, 8), arr(b, 8), p9)
        assert list(o9) == list(p9)
        assert val(o9) % R == a * b * RINV % R
