"""TEST INFRASTRUCTURE ONLY -- Python big-int restatement of the reference's multilinear KZG (kzg/src/multilinear_kzg.rs,
kzg/src/trusted_setup.rs, kzg/src/utils.rs) over BLS12-381 G1, the checker of zksc_g1_msm / zksc_kzg_open.

PARITY STATUS: the curve arithmetic lives in the third-party crates ark-ec / ark-test-curves 0.4.2 (Cargo.toml:21,32), absent from
/root/reference; it is restated here from the published BLS12-381 parameters (affine chord-and-tangent formulas on python ints --
deliberately a different algorithm from the device's Jacobian formulas) and pinned by: the generator lies on y^2 = x^3 + 4 and has
order r (tests/test_oracle_kzg.py).  The reference's own KZG tests (multilinear_kzg.rs:132-199) only assert `verify == true`;
`verify` is a pairing equation, restated here IN THE EXPONENT (the trusted setup's tau is known to a test), see `verify_in_exponent`:
an opening accepted by it is one the reference's pairing check accepts.  No reference test pins a commitment's coordinates:
"parity unpinned" at that level.  Each function cites the reference lines it follows (paths relative to /root/reference)."""
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
G1 = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
      0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
RQ = 1 << 384          # Montgomery radix of Fq in ark-ff (6 x u64)


def on_curve(pt):
    return pt is None or (pt[1] * pt[1] - pt[0] ** 3 - 4) % P == 0


def add(a, b):
    """affine group law; None is the identity"""
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def mul(k, pt):
    """pt.mul_bigint(k)"""
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


def boolean_hypercube(n):  # polynomial/src/utils.rs boolean_hypercube: rows in counting order, variable 0 first
    return [[(i >> (n - 1 - k)) & 1 for k in range(n)] for i in range(1 << n)]


def check_for_zero_and_one(bh, value):  # kzg/src/utils.rs:19-33
    acc = 1
    for b, e in zip(bh, value):
        acc = acc * ((1 - e) if b == 0 else e) % R
    return acc


def generate_array_of_points(n, eval_points):  # kzg/src/utils.rs:35-40: the Lagrange basis over {0,1}^n at tau
    return [check_for_zero_and_one(bh, eval_points) for bh in boolean_hypercube(n)]


class TrustedSetup:  # kzg/src/trusted_setup.rs:15-47 (G1 part; the G2 powers only enter the pairing check)
    def __init__(self, eval_points):
        self.tau = [t % R for t in eval_points]
        self.scalars = generate_array_of_points(len(eval_points), self.tau)
        self.powers_of_tau_in_g1 = [mul(s, G1) for s in self.scalars]

    setup = classmethod(lambda cls, pts: cls(pts))


def partial_evaluation0(ev, r):  # Multilinear::partial_evaluation(r, 0), evaluation_form.rs:123-141
    h = len(ev) // 2
    return [(ev[i] + r * (ev[i + h] - ev[i])) % R for i in range(h)]


def evaluation(ev, pts):  # evaluation_form.rs:162-175
    for r in pts:
        ev = partial_evaluation0(ev, r)
    return ev[0]


def commitment(ev, srs):  # multilinear_kzg.rs:33-48
    assert len(srs.powers_of_tau_in_g1) == len(ev)
    acc = None
    for c, pw in zip(ev, srs.powers_of_tau_in_g1):
        acc = add(acc, mul(c, pw))
    return acc


def get_poly_quotient(ev):  # kzg/src/utils.rs:12-17: f(1, .) - f(0, .)
    h = len(ev) // 2
    return [(ev[h + i] - ev[i]) % R for i in range(h)]


def add_to_front(ev, variable_length):  # evaluation_form.rs:86-96
    return (ev + ev) * (1 << variable_length)


def blown_polys(ev, points):
    """the polynomials MultilinearKZG::open commits to, one per variable (multilinear_kzg.rs:58-80)"""
    out, poly, n = [], list(ev), len(points)
    for idx, z in enumerate(points):
        q = get_poly_quotient(poly)
        if idx != n - 1:
            out.append(add_to_front(q, idx))
            poly = partial_evaluation0(poly, z)              # get_poly_remainder
        else:
            out.append(add_to_front(q + q, idx - 1) if idx >= 1 else q + q)      # duplicate_evaluation, then add_to_front(idx - 1)
            poly = partial_evaluation0(poly, z)
    return out, poly[0]


def open_(ev, points, srs):  # multilinear_kzg.rs:50-88 -> (evaluation, proofs)
    blown, final = blown_polys(ev, points)
    v = evaluation(ev, points)
    assert v == final, "Evaluation and final remainder mismatch!"      # :84-86
    return v, [commitment(b, srs) for b in blown]


def verify_in_exponent(ev, points, srs):
    """MultilinearKZG::verify (multilinear_kzg.rs:90-116) is  e(C - v G1, G2) == sum_i e(proof_i, tau_i G2 - z_i G2).  With
    C = f(tau) G1 and proof_i = q_i(tau) G1 this is the scalar identity  f(tau) - v == sum_i q_i(tau) (tau_i - z_i)  (mod r),
    checked here with the discrete logs the oracle knows: it holds iff the pairing check passes."""
    blown, _ = blown_polys(ev, points)
    f_tau = sum(c * s for c, s in zip(ev, srs.scalars)) % R
    v = evaluation(ev, points)
    rhs = sum((sum(c * s for c, s in zip(b, srs.scalars)) % R) * (t - z) for b, t, z in zip(blown, srs.tau, points)) % R
    return (f_tau - v) % R == rhs


def verify_group(commit, points, opening, srs):
    """MultilinearKZG::verify (multilinear_kzg.rs:90-116) on the group elements:  e(C - v G1, G2) == sum_i e(proof_i, (tau_i - z_i) G2)
    holds iff  C - v G1 == sum_i (tau_i - z_i) proof_i  (bilinearity; tau is known to a test's trusted setup)."""
    v, proofs = opening
    lhs = add(commit, mul(R - v % R, G1))
    rhs = None
    for pr, t, z in zip(proofs, srs.tau, points):
        rhs = add(rhs, mul((t - z) % R, pr))
    return len(proofs) == len(srs.tau) and lhs == rhs


# ---- ark-ec memory form: Projective { x, y, z } Jacobian, each Fq 6 x u64 little-endian, Montgomery (R = 2^384) ----
def to_ark(pt):
    import numpy as np
    if pt is None:
        x, y, z = RQ % P, RQ % P, 0
    else:
        x, y, z = pt[0] * RQ % P, pt[1] * RQ % P, RQ % P
    return np.frombuffer(b"".join(v.to_bytes(48, "little") for v in (x, y, z)), dtype=np.uint64).copy()


def from_ark(arr):
    b = bytes(memoryview(arr).cast("B"))
    rinv = pow(RQ, -1, P)
    x, y, z = (int.from_bytes(b[48 * i:48 * i + 48], "little") * rinv % P for i in range(3))
    if z == 0:
        return None
    zi = pow(z, -1, P)
    return (x * zi * zi % P, y * zi * zi * zi % P)
