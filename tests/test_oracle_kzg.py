"""The KZG oracle (oracle/kzgmodel.py) on the CPU: curve constants, group law, the reference's two KZG test inputs
(kzg/src/multilinear_kzg.rs:132-199) through the verifier's equation in the exponent, the helper known answers of kzg/src/utils.rs."""
import random

from oracle import kzgmodel as k

R = k.R


def test_curve_constants_and_group_law():
    assert k.on_curve(k.G1)
    assert k.mul(R, k.G1) is None                       # the generator has order r
    assert k.mul(R - 1, k.G1) == (k.G1[0], (-k.G1[1]) % k.P)
    a, b = k.mul(123456789, k.G1), k.mul(987654321, k.G1)
    assert k.on_curve(a) and k.add(a, b) == k.mul(123456789 + 987654321, k.G1)
    assert k.add(a, a) == k.mul(2 * 123456789, k.G1)
    assert k.from_ark(k.to_ark(a)) == a and k.from_ark(k.to_ark(None)) is None


def test_utils_known_answers():
    """the known answers of kzg/src/utils.rs:72-198"""
    F = lambda v: v % R
    ev = [0, 7, 0, 5, 0, 7, 4, 9]
    assert k.get_poly_quotient(ev) == [0, 0, 4, 4]                                           # test_get_poly_quotient :106-131
    assert k.get_poly_quotient([0, 7, 20, 25]) == [20, 18]
    assert k.get_poly_quotient([180, 169]) == [F(-11)]
    assert k.partial_evaluation0(ev, 5) == [0, 7, 20, 25]                                    # test_get_poly_remainder :134-162
    assert k.partial_evaluation0([0, 7, 20, 25], 9) == [180, 169]
    assert k.partial_evaluation0([180, 169], 6) == [114]
    want = [F(v) for v in (-6, 8, 9, -12, 12, -16, -18, 24)]
    assert [k.check_for_zero_and_one(bh, [2, 3, 4]) for bh in k.boolean_hypercube(3)] == want   # test_check_for_zero_and_one :73-103
    assert k.generate_array_of_points(3, [2, 3, 4]) == want                                  # test_generate_array_of_points :165-188


def test_reference_kzg_cases_verify_in_the_exponent():
    """test_kzg_1 / test_kzg_2: commitment + open, then the verifier's equation (restated on the discrete logs); a tampered setup fails"""
    for prover, verifier, ev in (([2, 3, 4], [5, 9, 6], [0, 7, 0, 5, 0, 7, 4, 9]),
                                 ([12, 9, 28, 40], [54, 90, 76, 160], [0, 0, 0, 2, 0, 0, 10, 12, 0, -12, 4, -6, 0, -12, 14, 4])):
        ev = [v % R for v in ev]
        srs = k.TrustedSetup(prover)
        assert k.verify_in_exponent(ev, verifier, srs)
        v, proofs = k.open_(ev, verifier, srs)
        c = k.commitment(ev, srs)
        # the same equation on the group elements: C - v G == sum_i (tau_i - z_i) proof_i
        lhs = k.add(c, k.mul(R - v, k.G1))
        rhs = None
        for pr, t, z in zip(proofs, srs.tau, verifier):
            rhs = k.add(rhs, k.mul((t - z) % R, pr))
        assert lhs == rhs
        bad = k.TrustedSetup([prover[0], prover[1] + 10] + prover[2:])
        rhs_bad = None
        for pr, t, z in zip(proofs, bad.tau, verifier):
            rhs_bad = k.add(rhs_bad, k.mul((t - z) % R, pr))
        assert lhs != rhs_bad                                                                 # tampered_tau_verify_status == false


def test_open_on_random_polynomials():
    rng = random.Random(9)
    for n in (2, 3, 5):
        ev = [rng.randrange(R) for _ in range(1 << n)]
        srs = k.TrustedSetup([rng.randrange(R) for _ in range(n)])
        assert k.verify_in_exponent(ev, [rng.randrange(R) for _ in range(n)], srs)


def test_succint_gkr_oracle_roundtrip():
    """gkr/src/succint_protocol.rs:281-351: the reference's two tests assert verify == true; here the oracle's own prover/verifier pair"""
    from oracle import gkrmodel as g
    c1 = g.Circuit([g.CircuitLayer([g.Gate("Mul", [0, 1])]), g.CircuitLayer([g.Gate("Add", [0, 1]), g.Gate("Mul", [2, 3])])])
    ev = c1.evaluation([2, 3, 4, 5])
    tau = k.TrustedSetup([54, 90])
    com, proof = g.SuccintGKRProtocol.prove(c1, ev, tau)
    assert g.SuccintGKRProtocol.verify(c1, com, proof, tau)
    assert not g.SuccintGKRProtocol.verify(c1, k.add(com, k.G1), proof, tau)
    bad = g.SuccintGKRProof(proof, (proof.proof_wb_opening[0] + 1, proof.proof_wb_opening[1]), proof.proof_wc_opening)
    assert not g.SuccintGKRProtocol.verify(c1, com, bad, tau)
