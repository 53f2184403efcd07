"""zksc_g1_msm / zksc_kzg_open (MultilinearKZG::commitment / open, kzg/src/multilinear_kzg.rs:33-88) against the big-int oracle:
the commitment and every opening proof must be the oracle's group element (compared in affine coordinates, where a point has one
representation), on the reference's own test inputs (:132-199) and on random polynomials; the device's Jacobian arithmetic must
handle the identity, doublings and cancelling points."""
import random

import numpy as np
import pytest

import zk_cryptography_b200 as zk
from zk_cryptography_b200.kzg import MultilinearKZG, TrustedSetup
from oracle import kzgmodel as k

pytestmark = pytest.mark.gpu
R = k.R


@pytest.fixture(scope="module", autouse=True)
def _ctx(ctx):
    return ctx


def _srs(model):
    return TrustedSetup(np.stack([k.to_ark(p) for p in model.powers_of_tau_in_g1]))


def _msm(ctx, scalars, points):
    out = np.zeros(18, dtype=np.uint64)
    s = zk.to_mont([int(v) % R for v in scalars]).reshape(-1, 4)
    pts = np.stack([k.to_ark(p) for p in points])
    ctx.check(zk.lib().zksc_g1_msm(ctx._h, zk._lib.p64(s), zk._lib.p64(pts), len(scalars), zk._lib.p64(out)))
    return k.from_ark(out)


def test_g1_arithmetic_special_cases(ctx):
    g, g2, g5 = k.G1, k.mul(2, k.G1), k.mul(5, k.G1)
    neg = (g[0], (-g[1]) % k.P)
    assert _msm(ctx, [1], [g]) == g
    assert _msm(ctx, [0], [g]) is None                                   # identity
    assert _msm(ctx, [R - 1], [g]) == neg
    assert _msm(ctx, [1, 1], [g, g]) == g2                               # equal points in one bucket: the doubling branch of the addition
    assert _msm(ctx, [1, 1], [g, neg]) is None                           # opposite points cancel
    assert _msm(ctx, [3, 2], [g, None]) == k.mul(3, g)                   # a point at infinity among the inputs
    assert _msm(ctx, [2, 3], [g5, g2]) == k.mul(16, g)
    assert _msm(ctx, [1 << 200, (1 << 254) + 12345], [g, g5]) == k.add(k.mul(1 << 200, g), k.mul((1 << 254) + 12345, g5))
    rng = random.Random(4)
    sc = [rng.randrange(R) for _ in range(300)]                           # more points than buckets of a window see on average
    pts = [k.mul(rng.randrange(1, 1 << 40), g) for _ in range(300)]
    want = None
    for s, p in zip(sc, pts):
        want = k.add(want, k.mul(s, p))
    assert _msm(ctx, sc, pts) == want


@pytest.mark.parametrize("prover,verifier,ev", [
    ([2, 3, 4], [5, 9, 6], [0, 7, 0, 5, 0, 7, 4, 9]),                                                          # test_kzg_1
    ([12, 9, 28, 40], [54, 90, 76, 160], [0, 0, 0, 2, 0, 0, 10, 12, 0, -12, 4, -6, 0, -12, 14, 4]),            # test_kzg_2
])
def test_reference_kzg_cases(ctx, prover, verifier, ev):
    ev = [v % R for v in ev]
    model = k.TrustedSetup(prover)
    srs = _srs(model)
    poly = zk.Multilinear(ev)
    assert k.from_ark(MultilinearKZG.commitment(poly, srs)) == k.commitment(ev, model)
    proof = MultilinearKZG.open(poly, verifier, srs)
    want_v, want_proofs = k.open_(ev, verifier, model)
    assert zk.from_mont(proof.evaluation) == want_v
    assert [k.from_ark(p) for p in proof.proofs] == want_proofs
    assert k.verify_in_exponent(ev, verifier, model)       # ... and those are openings the reference's pairing check accepts


def test_random_polynomials(ctx):
    rng = random.Random(11)
    for n in (1, 2, 5, 7):
        ev = [rng.randrange(R) for _ in range(1 << n)]
        model = k.TrustedSetup([rng.randrange(R) for _ in range(n)])
        srs = _srs(model)
        z = [rng.randrange(R) for _ in range(n)]
        poly = zk.Multilinear(ev)
        assert k.from_ark(MultilinearKZG.commitment(poly, srs)) == k.commitment(ev, model)
        if n >= 2:
            proof = MultilinearKZG.open(poly, z, srs)
            want_v, want_proofs = k.open_(ev, z, model)
            assert zk.from_mont(proof.evaluation) == want_v and [k.from_ark(p) for p in proof.proofs] == want_proofs
    with pytest.raises(zk.ZkscError):
        MultilinearKZG.commitment(zk.Multilinear([1, 2, 3, 4]), _srs(k.TrustedSetup([3])))      # lengths must tally (:36-41)
