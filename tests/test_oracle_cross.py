"""The two oracles (C restatement, Python big-int model) must agree byte for byte on random seeds, for all
four provers, including structured inputs that exercise the sparse rules."""
import random

import pytest

from oracle import cref
from oracle import pymodel as pm

R = pm.R_MOD


def _gen(kind, n, rng):
    if kind == "rand":
        return [rng.randrange(R) for _ in range(1 << n)]
    if kind == "small":
        return [rng.randrange(256) for _ in range(1 << n)]
    if kind == "zero":
        return [0] * (1 << n)
    if kind == "ones":
        return [1] * (1 << n)
    if kind == "ramp":  # the reference's bench inputs: 0..=255 (sumcheck_benchmark.rs:7-12)
        return [i % 256 for i in range(1 << n)]
    return [rng.randrange(2) for _ in range(1 << n)]


def pm_prove(proto, polys, s):
    if proto == 0:
        sc = pm.Sumcheck(polys[0].polys[0])
        sc.sum = s
        pr, ch = sc.prove()
        return b"".join(u.to_bytes() for u in pr.univariate_poly), ch
    if proto == 1:
        pr, ch = pm.ComposedSumcheck(polys[0]).prove()
        return b"".join(pm.vec_to_bytes(rp) for rp in pr.round_polys), ch
    if proto == 2:
        pr, ch = pm.MultiComposedSumcheckProver.prove_partial(polys, s)
        return pr.to_bytes(), ch
    pr, ch = pm.MultiComposedSumcheckProver.prove(polys, s)
    return pr.to_bytes(), ch


@pytest.mark.parametrize("seed", range(24))
def test_c_oracle_equals_python_model(seed):
    rng = random.Random(1000 + seed)
    n = rng.randint(1, 7)
    proto = seed % 4
    degs = [rng.randint(1, 5) for _ in range(rng.randint(1, 3))]
    if proto == 0:
        degs = [1]
    if proto == 1:
        degs = degs[:1]
    kind = ["rand", "small", "zero", "bits", "ones", "ramp"][seed % 6]
    tabs = [[_gen(kind, n, rng) for _ in range(d)] for d in degs]
    polys = [pm.ComposedMultilinear([pm.Multilinear(t) for t in tp]) for tp in tabs]
    flat = cref.ints_to_canon([v for tp in tabs for t in tp for v in t])
    s = pm.MultiComposedSumcheckProver.calculate_poly_sum(polys)
    assert s == cref.poly_sum(n, degs, flat)
    if seed % 5 == 0:
        s = (s + 1) % R  # the prover does not recompute the claimed sum (multi_composed_sumcheck.rs:70)
    want = pm_prove(proto, polys, s)
    assert cref.prove(proto, n, degs, flat, s) == want


def test_threads_do_not_change_results():
    rng = random.Random(5)
    n, degs = 13, [2, 3]
    flat = cref.ints_to_canon([rng.randrange(R) for _ in range(sum(degs) << n)])
    s = cref.poly_sum(n, degs, flat)
    cref.set_threads(1)
    a = cref.prove(2, n, degs, flat, s)
    cref.set_threads(4)
    b = cref.prove(2, n, degs, flat, s)
    cref.set_threads(1)
    assert a == b


def test_synth_tables_agree():
    for seed, tab, n in ((7, 1, 3), (2026, 0, 6), (2**63 + 5, 4, 4)):
        assert cref.canon_to_ints(cref.synth_table(seed, tab, n)) == pm.synth_table(seed, tab, n).evaluations


def test_verify_partial_c_vs_python():
    rng = random.Random(9)
    polys = [pm.ComposedMultilinear([pm.Multilinear([rng.randrange(R) for _ in range(16)]) for _ in range(2)]) for _ in range(2)]
    s = pm.MultiComposedSumcheckProver.calculate_poly_sum(polys)
    pr, ch = pm.MultiComposedSumcheckProver.prove_partial(polys, s)
    sub = pm.MultiComposedSumcheckVerifier.verify_partial(pr)
    ok, csum, cch = cref.verify_partial(4, s, [rp.monomial for rp in pr.round_polys])
    assert ok and csum == sub.sum and cch == sub.challenges == ch
    ok, _, _ = cref.verify_partial(4, (s + 1) % R, [rp.monomial for rp in pr.round_polys])
    assert not ok
