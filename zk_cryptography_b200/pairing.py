"""BLS12-381 pairing on the host: the verifier half of the reference's multilinear KZG (MultilinearKZG::verify,
kzg/src/multilinear_kzg.rs:90-116; sum_pairing_results, kzg/src/utils.rs:42-61) and of SuccintGKRProtocol::verify.

The reference gets the pairing from ark-ec / ark-test-curves 0.4.2 (`Bls12_381::pairing`, third-party, absent from /root/reference).
A verifier computes n + 1 pairings of single points: host-side, constant-size work, like the Fiat-Shamir transcript -- nothing here is
table-sized, so nothing here runs on the GPU.  Plain Python integers, textbook algorithms chosen for being easy to check rather than
fast (~0.15 s per pairing):

  Fq2 = Fq[u]/(u^2 + 1),  Fq6 = Fq2[v]/(v^3 - (1 + u)),  Fq12 = Fq6[w]/(w^2 - v)            (the tower ark-ff uses for this curve)
  G1: y^2 = x^3 + 4 over Fq;  G2: y^2 = x^3 + 4 (1 + u) over Fq2, mapped into E(Fq12) by (x, y) -> (x / w^2, y / w^3)   (w^6 = 1 + u)
  e(P, Q) = f_{|x|, Q}(P) ^ ((p^12 - 1) / r)  with the affine Miller loop over E(Fq12), x = -0xd201000000010000 (f conjugated for
  the sign); vertical lines are dropped (they lie in Fq6 and die in the final exponentiation).

Checked on the CPU (tests/test_host_pairing.py): both generators lie on their curves and have order r, e is bilinear and
non-degenerate with values of order r, and the reference's own KZG tests (`verify == true`, tampered setup `== false`,
multilinear_kzg.rs:132-199) hold.
"""
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X_ABS = 0xd201000000010000          # |x| of the BLS12-381 parameter x = -X_ABS
G1 = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
      0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
G2 = ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
       0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
      (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
       0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))

# ---- Fq2: (a, b) = a + b u, u^2 = -1 ------------------------------------------------------------------------------------
F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (1, 1)                          # the non-residue 1 + u


def f2_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def f2_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def f2_neg(x): return ((-x[0]) % P, (-x[1]) % P)
def f2_mul(x, y): return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def f2_muls(x, k): return (x[0] * k % P, x[1] * k % P)


def f2_inv(x):
    d = pow(x[0] * x[0] + x[1] * x[1], -1, P)
    return (x[0] * d % P, (-x[1]) * d % P)


# ---- Fq6: (c0, c1, c2) = c0 + c1 v + c2 v^2, v^3 = XI --------------------------------------------------------------------
F6_ZERO, F6_ONE = (F2_ZERO, F2_ZERO, F2_ZERO), (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(x, y): return tuple(f2_add(a, b) for a, b in zip(x, y))
def f6_sub(x, y): return tuple(f2_sub(a, b) for a, b in zip(x, y))
def f6_neg(x): return tuple(f2_neg(a) for a in x)


def f6_mul(x, y):
    a0, a1, a2 = x
    b0, b1, b2 = y
    t00, t11, t22 = f2_mul(a0, b0), f2_mul(a1, b1), f2_mul(a2, b2)
    c0 = f2_add(t00, f2_mul(XI, f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul(XI, t22))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a2, b0)), t11)
    return (c0, c1, c2)


def f6_mul_by_v(x):                  # (c0 + c1 v + c2 v^2) v = XI c2 + c0 v + c1 v^2
    return (f2_mul(XI, x[2]), x[0], x[1])


def f6_inv(x):
    a0, a1, a2 = x
    t0 = f2_sub(f2_mul(a0, a0), f2_mul(XI, f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul(XI, f2_mul(a2, a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_mul(a1, a1), f2_mul(a0, a2))
    d = f2_inv(f2_add(f2_mul(a0, t0), f2_mul(XI, f2_add(f2_mul(a2, t1), f2_mul(a1, t2)))))
    return (f2_mul(t0, d), f2_mul(t1, d), f2_mul(t2, d))


# ---- Fq12: (c0, c1) = c0 + c1 w, w^2 = v ----------------------------------------------------------------------------------
F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(x, y):
    a0, a1 = x
    b0, b1 = y
    t0, t1 = f6_mul(a0, b0), f6_mul(a1, b1)
    return (f6_add(t0, f6_mul_by_v(t1)), f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(b0, b1)), t0), t1))


def f12_sub(x, y): return (f6_sub(x[0], y[0]), f6_sub(x[1], y[1]))
def f12_add(x, y): return (f6_add(x[0], y[0]), f6_add(x[1], y[1]))
def f12_conj(x): return (x[0], f6_neg(x[1]))        # the p^6 Frobenius


def f12_inv(x):
    a0, a1 = x
    d = f6_inv(f6_sub(f6_mul(a0, a0), f6_mul_by_v(f6_mul(a1, a1))))
    return (f6_mul(a0, d), f6_neg(f6_mul(a1, d)))


def f12_pow(x, e):
    acc = F12_ONE
    for bit in bin(e)[2:]:
        acc = f12_mul(acc, acc)
        if bit == "1":
            acc = f12_mul(acc, x)
    return acc


def f12_from_fq(a): return (((a % P, 0), F2_ZERO, F2_ZERO), F6_ZERO)
def f12_from_fq2(a): return ((a, F2_ZERO, F2_ZERO), F6_ZERO)


W = (F6_ZERO, F6_ONE)                                # w
W2_INV = f12_inv(f12_mul(W, W))                       # 1 / w^2
W3_INV = f12_inv(f12_mul(f12_mul(W, W), W))           # 1 / w^3


# ---- groups (affine; None is the identity) ---------------------------------------------------------------------------------
def _ec_add(a, b, add, sub, mul, inv, dbl_num):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if add(y1, y2) in (0, F2_ZERO):
            return None
        lam = mul(dbl_num(x1), inv(add(y1, y1)))
    else:
        lam = mul(sub(y2, y1), inv(sub(x2, x1)))
    x3 = sub(sub(mul(lam, lam), x1), x2)
    return (x3, sub(mul(lam, sub(x1, x3)), y1))


def g1_add(a, b):
    return _ec_add(a, b, lambda x, y: (x + y) % P, lambda x, y: (x - y) % P, lambda x, y: x * y % P, lambda x: pow(x, -1, P), lambda x: 3 * x * x % P)


def g2_add(a, b):
    return _ec_add(a, b, f2_add, f2_sub, f2_mul, f2_inv, lambda x: f2_muls(f2_mul(x, x), 3))


def _ec_mul(k, pt, add):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


def g1_mul(k, pt): return _ec_mul(k, pt, g1_add)
def g2_mul(k, pt): return _ec_mul(k, pt, g2_add)
def g1_neg(pt): return None if pt is None else (pt[0], (-pt[1]) % P)
def g2_neg(pt): return None if pt is None else (pt[0], f2_neg(pt[1]))
def g1_on_curve(pt): return pt is None or (pt[1] * pt[1] - pt[0] ** 3 - 4) % P == 0
def g2_on_curve(pt): return pt is None or f2_sub(f2_mul(pt[1], pt[1]), f2_add(f2_mul(f2_mul(pt[0], pt[0]), pt[0]), f2_muls(XI, 4))) == F2_ZERO


# ---- the pairing --------------------------------------------------------------------------------------------------------------
FINAL_EXP = (P ** 12 - 1) // R


def miller_loop(p1, q2):
    """f_{|x|, psi(Q)}(P), conjugated for x < 0; 1 when either point is the identity"""
    if p1 is None or q2 is None:
        return F12_ONE
    xp, yp = f12_from_fq(p1[0]), f12_from_fq(p1[1])
    qx, qy = f12_mul(f12_from_fq2(q2[0]), W2_INV), f12_mul(f12_from_fq2(q2[1]), W3_INV)      # psi(Q) on E(Fq12): y^2 = x^3 + 4
    tx, ty = qx, qy
    f = F12_ONE
    three = f12_from_fq(3)
    for bit in bin(X_ABS)[3:]:
        lam = f12_mul(f12_mul(three, f12_mul(tx, tx)), f12_inv(f12_add(ty, ty)))              # tangent at T
        f = f12_mul(f12_mul(f, f), f12_sub(f12_sub(yp, ty), f12_mul(lam, f12_sub(xp, tx))))
        nx = f12_sub(f12_sub(f12_mul(lam, lam), tx), tx)
        tx, ty = nx, f12_sub(f12_mul(lam, f12_sub(tx, nx)), ty)
        if bit == "1":
            lam = f12_mul(f12_sub(qy, ty), f12_inv(f12_sub(qx, tx)))                          # chord through T and Q
            f = f12_mul(f, f12_sub(f12_sub(yp, ty), f12_mul(lam, f12_sub(xp, tx))))
            nx = f12_sub(f12_sub(f12_mul(lam, lam), tx), qx)
            tx, ty = nx, f12_sub(f12_mul(lam, f12_sub(tx, nx)), ty)
    return f12_conj(f)


def final_exponentiation(f):
    return f12_pow(f, FINAL_EXP)


def pairing(p1, q2):
    """e(P, Q) for affine P in G1 (pair of ints) and Q in G2 (pair of Fq2); an element of Fq12 of order r"""
    return final_exponentiation(miller_loop(p1, q2))


def multi_pairing(pairs):
    """prod_i e(P_i, Q_i) with one final exponentiation"""
    f = F12_ONE
    for p1, q2 in pairs:
        f = f12_mul(f, miller_loop(p1, q2))
    return final_exponentiation(f)


# ---- ark-ec memory forms ------------------------------------------------------------------------------------------------------
RQ = 1 << 384


def g1_from_ark(arr):
    """(18,) uint64 Jacobian / Montgomery (ark-ec G1Projective) -> affine ints"""
    b = bytes(memoryview(arr).cast("B"))
    rinv = pow(RQ, -1, P)
    x, y, z = (int.from_bytes(b[48 * i:48 * i + 48], "little") * rinv % P for i in range(3))
    if z == 0:
        return None
    zi = pow(z, -1, P)
    return (x * zi * zi % P, y * zi * zi * zi % P)


# ---- the native form (csrc/host_pairing.hpp through the C ABI): same construction, ~40 ms per verifier check instead of seconds -----------
def _limbs(x, n=6):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def _pack_g1(pt):
    return [0] * 12 if pt is None else _limbs(pt[0]) + _limbs(pt[1])


def _pack_g2(pt):
    return [0] * 24 if pt is None else _limbs(pt[0][0]) + _limbs(pt[0][1]) + _limbs(pt[1][0]) + _limbs(pt[1][1])


def native_pairing_check(pairs):
    """prod_i e(P_i, Q_i) == 1, computed by libzksc.so on the host (zksc_pairing_check; no GPU involved)"""
    import ctypes

    import numpy as np

    from . import _lib
    g1 = np.array([w for p1, _ in pairs for w in _pack_g1(p1)], dtype=np.uint64)
    g2 = np.array([w for _, q2 in pairs for w in _pack_g2(q2)], dtype=np.uint64)
    ok = ctypes.c_int(0)
    rc = _lib.lib().zksc_pairing_check(_lib.p64(g1), _lib.p64(g2), len(pairs), ctypes.byref(ok))
    if rc != 0:
        raise _lib.ZkscError(rc, "zksc_pairing_check: a coordinate is not reduced or a point is not on its curve")
    return bool(ok.value)


def native_pairing(p1, q2):
    """e(P, Q) from the native code, in this module's tuple form"""
    import numpy as np

    from . import _lib
    out = np.zeros(72, dtype=np.uint64)
    rc = _lib.lib().zksc_pairing(_lib.p64(np.array(_pack_g1(p1), dtype=np.uint64)), _lib.p64(np.array(_pack_g2(q2), dtype=np.uint64)), _lib.p64(out))
    if rc != 0:
        raise _lib.ZkscError(rc, "zksc_pairing: a coordinate is not reduced or a point is not on its curve")
    c = [sum(int(out[6 * i + j]) << (64 * j) for j in range(6)) for i in range(12)]
    f2 = [(c[2 * i], c[2 * i + 1]) for i in range(6)]
    return ((f2[0], f2[1], f2[2]), (f2[3], f2[4], f2[5]))
