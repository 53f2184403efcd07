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
    from zk_cryptography_b200 import pairing as pr
    return TrustedSetup(np.stack([k.to_ark(p) for p in model.powers_of_tau_in_g1]), [pr.g2_mul(t, pr.G2) for t in model.tau])


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
    # MultilinearKZG::verify itself (multilinear_kzg.rs:90-116: the pairing equation, host side) on what the GPU produced -- the
    # reference's own assertions: true, and false for a setup tampered with (:195-198)
    commit = MultilinearKZG.commitment(poly, srs)
    assert MultilinearKZG.verify(commit, verifier, proof, srs) is True
    bad = list(prover)
    bad[1] += 10
    assert MultilinearKZG.verify(commit, verifier, proof, _srs(k.TrustedSetup(bad))) is False


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


# ---- SuccintGKRProtocol::prove (gkr/src/succint_protocol.rs:37-167) on top of zksc_gkr_prove + the KZG entry points ----
SUCCINT_1 = ([[("Mul", [0, 1])], [("Add", [0, 1]), ("Mul", [2, 3])]], [2, 3, 4, 5], [54, 90])                                   # :281-306
SUCCINT_2 = ([[("Add", [0, 1])], [("Mul", [0, 1]), ("Add", [2, 3])],                                                            # :309-351
              [("Add", [0, 1]), ("Mul", [2, 3]), ("Mul", [4, 5]), ("Mul", [6, 7])]], [4, 3, 7, 6, 6, 1, 4, 2], [54, 90, 76])


@pytest.mark.parametrize("layers,inp,points", [SUCCINT_1, SUCCINT_2, (SUCCINT_2[0], SUCCINT_2[1], [54, 90, 76, 11])])
def test_succint_gkr_prove(ctx, layers, inp, points):
    """the reference's two succinct-GKR tests (and one with a trusted setup larger than the input layer: the add_to_back blow-up):
    commitment, both openings and every sumcheck byte equal the oracle's, and the oracle's verifier accepts"""
    from oracle import gkrmodel as g
    zc = zk.Circuit([zk.CircuitLayer([zk.Gate(zk.GateType.Add if t == "Add" else zk.GateType.Mul, i) for t, i in layer]) for layer in layers])
    oc = g.Circuit([g.CircuitLayer([g.Gate(t, i) for t, i in layer]) for layer in layers])
    ev = zc.evaluation(inp)
    if layers is SUCCINT_2[0]:
        assert ev[0][0] == 308                                                            # :337
    model = k.TrustedSetup(points)
    commitment, proof = zk.SuccintGKRProtocol.prove(zc, ev, _srs(model))
    want_c, want = g.SuccintGKRProtocol.prove(oc, ev, model)
    assert k.from_ark(commitment) == want_c
    assert b"".join(p.to_bytes() for p in proof.sumcheck_proofs) == b"".join(p.to_bytes() for p in want.sumcheck_proofs)
    assert proof.wb_s == want.wb_s and proof.wc_s == want.wc_s
    for got, exp in ((proof.proof_wb_opening, want.proof_wb_opening), (proof.proof_wc_opening, want.proof_wc_opening)):
        assert zk.from_mont(got.evaluation) == exp[0] and [k.from_ark(p) for p in got.proofs] == exp[1]
    assert g.SuccintGKRProtocol.verify(oc, want_c, want, model)
    # the product's proof through the oracle's verifier
    got = g.SuccintGKRProof(want, (zk.from_mont(proof.proof_wb_opening.evaluation), [k.from_ark(p) for p in proof.proof_wb_opening.proofs]),
                            (zk.from_mont(proof.proof_wc_opening.evaluation), [k.from_ark(p) for p in proof.proof_wc_opening.proofs]))
    assert g.SuccintGKRProtocol.verify(oc, k.from_ark(commitment), got, model)
    assert not g.SuccintGKRProtocol.verify(oc, k.add(want_c, k.G1), want, model)          # a wrong commitment is rejected
    # SuccintGKRProtocol::verify (succint_protocol.rs:169-266) of the product, pairing checks included: the reference's assertion
    srs = _srs(model)
    assert zk.SuccintGKRProtocol.verify(zc, commitment, proof, srs) is True
    proof.wb_s[-1] = (proof.wb_s[-1] + 1) % R
    assert zk.SuccintGKRProtocol.verify(zc, commitment, proof, srs) is False
