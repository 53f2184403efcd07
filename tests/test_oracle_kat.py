"""The oracles against every primitive known-answer test the reference holds for this path
(SURVEY.md section 4 / 8c).  File:line references are to /root/reference."""
import hashlib

import pytest

from oracle import cref
from oracle import pymodel as pm

R = pm.R_MOD
F = pm.fr


def ML(v):
    return pm.Multilinear([F(x) for x in v])


def test_partial_evaluation_1():  # evaluation_form.rs:315-325
    assert ML([3, 1, 2, 5]).partial_evaluation(F(5), 0) == ML([-2, 21])
    assert cref.partial_evaluation([3, 1, 2, 5], 5, 0) == [F(-2), 21]


def test_partial_evaluation_2():  # evaluation_form.rs:328-358
    ev = [3, 9, 7, 13, 6, 12, 10, 18]
    p = ML(ev)
    assert p.partial_evaluation(2, 0).evaluation([3, 2]) == 57
    assert p.partial_evaluation(3, 1).evaluation([3, 2]) == 72
    assert p.partial_evaluation(1, 2).evaluation([3, 2]) == 38
    for r, k, want in ((2, 0, 57), (3, 1, 72), (1, 2, 38)):
        assert cref.evaluation(cref.partial_evaluation(ev, r, k), [3, 2]) == want


def test_evaluation():  # evaluation_form.rs:362-405
    assert ML([3, 1, 2, 5]).evaluation([5, 6]) == 136
    assert ML([3, 9, 7, 13, 6, 12, 10, 18]).evaluation([2, 3, 1]) == 39
    assert ML([0, 0, 0, 3, 0, 0, 2, 5]).evaluation([2, 3, 4]) == 48
    assert cref.evaluation([3, 1, 2, 5], [5, 6]) == 136
    assert cref.evaluation([3, 9, 7, 13, 6, 12, 10, 18], [2, 3, 1]) == 39
    assert cref.evaluation([0, 0, 0, 3, 0, 0, 2, 5], [2, 3, 4]) == 48


def test_half_sums():  # evaluation_form.rs:408-438
    assert ML([0, 0, 0, 2, 2, 2, 2, 4]).split_poly_into_two_and_sum_each_part() == ML([2, 10])
    assert ML([0, 0, 2, 7, 3, 3, 6, 11]).split_poly_into_two_and_sum_each_part() == ML([9, 23])
    assert ML([1, 2, 3, 4, 5, 6, 7, 8]).sum_over_the_boolean_hypercube() == 36  # :440-462


def test_add_mul_distinct():  # evaluation_form.rs:264-312
    a, b = ML([0, 0, 2, 2]), ML([0, 3, 0, 3])
    assert a.add_distinct(b) == ML([0, 3, 0, 3, 0, 3, 0, 3, 2, 5, 2, 5, 2, 5, 2, 5])
    assert a.mul_distinct(b) == ML([0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 0, 6, 0, 6, 0, 6])


def test_poly_subtraction():  # evaluation_form.rs:508-545
    assert ML([0, 0, 0, 5, 4, 4, 7, 12]).sub(ML([0, 0, 0, 2, 0, 0, 1, 3])) == ML([0, 0, 0, 3, 4, 4, 6, 9])


def test_element_wise_product_and_composed_eval():  # composed_multilinear.rs:134-170
    c = pm.ComposedMultilinear([ML([0, 1, 2, 3]), ML([0, 0, 0, 1])])
    assert c.element_wise_product() == [0, 0, 0, 3]
    # (2a+3)(ab) style evaluation = 42 : f1 = [0,0,0,... ] reference builds 2 polys evaluating to 42
    c2 = pm.ComposedMultilinear([ML([0, 0, 0, 2]), ML([0, 3, 0, 3])])
    assert c2.evaluation([2, 3]) == (ML([0, 0, 0, 2]).evaluation([2, 3]) * ML([0, 3, 0, 3]).evaluation([2, 3])) % R


def test_interpolation_kats():  # sparse_univariate.rs:361-446 (structural equality where the reference asserts it)
    S = pm.SparseUnivariatePolynomial
    p = S.interpolation([(F(1), F(2)), (F(2), F(3)), (F(4), F(11))])
    assert p == S.new([3, 0, F(-2), 1, 1, 2]) and p.evaluate(2) == 3
    p5 = S.interpolation([(F(i), F(y)) for i, y in ((1, 6), (2, 11), (3, 18), (4, 27), (5, 38))])
    assert p5 == S.new([3, 0, 2, 1, 1, 2]) and p5.evaluate(2) == 11
    assert S.interpolation([(0, 0), (1, 2)]).evaluate(2) == 4                      # 2x (zero constant dropped)
    assert S.interpolation([(0, 0), (1, 2)]).monomial == [(2, 1)]
    q = S.interpolation([(0, 5), (1, 7), (2, 13)])                                  # 2x^2 + 5 (zero x term dropped)
    assert q.evaluate(2) == 13 and q.monomial == [(5, 0), (2, 2)]
    assert S.interpolation([(0, 12), (1, 48), (3, 3150), (4, 11772), (5, 33452), (8, 315020)]).evaluate(1) == 48
    assert S.interpolation([(0, 0), (1, 5), (2, 14)]).evaluate(2) == 14
    # C oracle, x = 0..n-1 form
    assert cref.interpolate([0, 2]) == [(2, 1)]
    assert cref.interpolate([5, 7, 13]) == [(5, 0), (2, 2)]
    assert cref.interpolate([3, 6, 11, 18, 27]) == [(3, 0), (2, 1), (1, 2)]


def test_sparse_add_and_eval():  # sparse_univariate.rs:232-299
    S = pm.SparseUnivariatePolynomial
    assert S.new([5, 0, 2, 1, 4, 6]).evaluate(2) == 265
    assert S.new([5, 0]).add(S.new([2, 1])) == S.new([5, 0, 2, 1])
    assert S.new([5, 0, 5, 2]).add(S.new([2, 1, 2, 2])) == S.new([5, 0, 2, 1, 7, 2])
    # zero SUMS are kept by Add (source text :171-176)
    assert S.new([5, 1]).add(S.new([F(-5), 1])).monomial == [(0, 1)]


def test_be32():  # sumcheck/src/utils.rs:71-93
    assert pm.be32(1) == bytes(31) + b"\x01" and pm.be32(100) == bytes(31) + bytes([100])
    assert cref.be32(1) == bytes(31) + b"\x01" and cref.be32(100) == bytes(31) + bytes([100])
    assert pm.be32(90) != bytes(31) + bytes([10])


def test_round_poly_format():  # sumcheck/src/utils.rs:125-141
    assert pm.convert_round_poly_to_uni_poly_format([1, 1, 1, 1]) == [(0, 1), (1, 1), (2, 1), (3, 1)]


def test_hypercube_sums():  # sumcheck.rs:108-123, composed_sumcheck.rs:108-140, multi_composed_sumcheck.rs:195-214, utils.rs:145-165
    s = pm.Sumcheck(ML([0, 0, 0, 2, 2, 2, 2, 4]))
    s.poly_sum()
    assert s.sum == 12
    CS, CM = pm.ComposedSumcheck, pm.ComposedMultilinear
    assert CS.calculate_poly_sum(CM([ML([0, 1, 2, 3]), ML([0, 0, 0, 1])])) == 3
    assert CS.calculate_poly_sum(CM([ML([3, 3, 5, 5]), ML([0, 0, 0, 1])])) == 5
    assert CS.calculate_poly_sum(CM([ML([0, 1, 2, 3])])) == 6
    assert CS.calculate_poly_sum(CM([ML([0, 0, 0, 2, 2, 2, 2, 4])])) == 12
    MP = pm.MultiComposedSumcheckProver
    assert MP.calculate_poly_sum([CM([ML([0, 1, 2, 3])]), CM([ML([0, 0, 0, 1])])]) == 7
    assert MP.calculate_poly_sum([CM([ML([0, 0, 0, 2])]), CM([ML([0, 3, 0, 3])])]) == 8
    assert pm.sum_over_boolean_hypercube([CM([ML([1, 2, 3, 4, 5, 6, 7, 8])])]) == 36
    assert cref.poly_sum(2, [1, 1], cref.ints_to_canon([0, 1, 2, 3, 0, 0, 0, 1])) == 7


REF_SUMCHECK_INPUTS = [  # sumcheck.rs:126-202 / composed_sumcheck.rs:167-241
    [0, 0, 2, 7, 3, 3, 6, 11],
    [0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0],
    [1, 3, 5, 7, 2, 4, 6, 8, 3, 5, 7, 9, 4, 6, 8, 10],
]


@pytest.mark.parametrize("ev", REF_SUMCHECK_INPUTS)
def test_reference_protocol_tests_roundtrip(ev):
    """every protocol test of the reference asserts verify == true on these inputs"""
    sc = pm.Sumcheck(ML(ev))
    sc.poly_sum()
    proof, _ = sc.prove()
    assert sc.verify(proof)
    cs = pm.ComposedSumcheck(pm.ComposedMultilinear([ML(ev)]))
    p2, _ = cs.prove()
    assert cs.verify(p2, pm.ComposedSumcheck.calculate_poly_sum(p2.poly))


def test_reference_composed_product_roundtrip():  # composed_sumcheck.rs:143-164
    cs = pm.ComposedSumcheck(pm.ComposedMultilinear([ML([3, 3, 5, 5]), ML([0, 0, 0, 1])]))
    p, _ = cs.prove()
    assert cs.verify(p, pm.ComposedSumcheck.calculate_poly_sum(p.poly))


def _multi_cases():
    p1, p2 = ML([0, 0, 0, 2]), ML([0, 3, 0, 3])
    CM = pm.ComposedMultilinear
    yield [CM([p1]), CM([p2])]                       # multi_composed_sumcheck.rs:217-231
    yield [CM([p1]), CM([p2]), CM([p2])]             # :233-248
    yield [CM([p1, p2]), CM([p2, p1])]               # :250-264
    add_i, mul_i = ML([4, 4, 7, 7, 4, 4, 7, 9]), ML([3, 3, 3, 4, 3, 3, 5, 6])
    w_b, w_c = ML([0, 4]), ML([0, 3])
    yield [CM([add_i.partial_evaluation(2, 0), w_b.add_distinct(w_c)]), CM([mul_i.partial_evaluation(2, 0), w_b.mul_distinct(w_c)])]  # :267-311


@pytest.mark.parametrize("idx", range(4))
def test_reference_multi_composed_roundtrip(idx):
    polys = list(_multi_cases())[idx]
    s = pm.MultiComposedSumcheckProver.calculate_poly_sum(polys)
    proof, ch = pm.MultiComposedSumcheckProver.prove(polys, s)
    assert pm.MultiComposedSumcheckVerifier.verify(polys, proof)
    pp, _ = pm.MultiComposedSumcheckProver.prove_partial(polys, s)
    sub = pm.MultiComposedSumcheckVerifier.verify_partial(pp)
    assert sub.sum == sum(p.evaluation(sub.challenges) for p in polys) % R


def test_derived_check_values_gkr_example():
    """SURVEY.md 8c 'derived check values' (derived by the survey's throw-away model, re-derived here
    independently; NOT emitted by the Rust binary): sum 213, round-0 poly 12 + 89x + 100x^2, challenges,
    SHA-256 of proof.to_bytes()."""
    polys = list(_multi_cases())[3]
    s = pm.MultiComposedSumcheckProver.calculate_poly_sum(polys)
    proof, ch = pm.MultiComposedSumcheckProver.prove(polys, s)
    assert s == 213 and proof.round_polys[0].monomial == [(12, 0), (89, 1), (100, 2)]
    assert ch[0] == 0x716D56C92CA17EBB13DC4AEF7135E1BCDB9FA3005A89F37E8B9A2D9C9BB061E4
    assert ch[1] == 0x7366CB6F4E2313524D45F834BF31E049F5F31A9716B2585D8E25D1B69EB8C02D
    assert hashlib.sha256(proof.to_bytes()).hexdigest() == "92e6503128821cd2e514d9c90ba198170b40f403e23b5757dc4f38e39c6dbbf2"
    sc = pm.Sumcheck(ML([0, 0, 2, 7, 3, 3, 6, 11]))
    sc.poly_sum()
    pr, _ = sc.prove()
    assert sc.sum == 32 and pr.univariate_poly[0].evaluations == [9, 23]


def test_transcript_semantics():  # fiat_shamir.rs:17-29: digest is fed back after each challenge
    t = pm.FiatShamirTranscript()
    t.commit(b"abc")
    c1, c2 = t.challenge(), t.challenge()
    assert c1 == hashlib.sha256(b"abc").digest() and c2 == hashlib.sha256(c1).digest()
    assert cref.transcript_two_challenges(b"abc") == (c1, c2)
    assert pm.from_be_bytes_mod_order(b"\xff" * 32) == (2**256 - 1) % R
