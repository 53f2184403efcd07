"""Parity of the CUDA path (through the C ABI) against the oracles: bit-exact field values, round
polynomials, challenges and proof bytes.  File:line references are to /root/reference."""
import random

import numpy as np
import pytest

import zk_cryptography_b200 as zk
from zk_cryptography_b200 import _lib
from oracle import cref
from oracle import pymodel as pm

pytestmark = pytest.mark.gpu
R = pm.R_MOD
F = pm.fr


@pytest.fixture(scope="module", autouse=True)
def _ctx(ctx):
    return ctx


def ML(v):
    return zk.Multilinear([F(x) for x in v])


# ---- device field arithmetic ---------------------------------------------------------------------
def test_device_field_ops_bit_exact(ctx):
    rng = random.Random(21)
    edge = [0, 1, 2, R - 1, R - 2, R // 2, (R + 1) // 2, 2**255 % R, (1 << 32) - 1, (1 << 64) - 1, R - (1 << 32)]
    a = edge + [rng.randrange(R) for _ in range(5000)]
    b = list(reversed(edge)) + [rng.randrange(R) for _ in range(5000)]
    n = 1 << 13
    a, b = (a * 2)[:n], (b * 2)[:n]
    A, B = zk.Multilinear(a), zk.Multilinear(b)
    assert (A + B).to_ints() == [(x + y) % R for x, y in zip(a, b)]
    assert (A - B).to_ints() == [(x - y) % R for x, y in zip(a, b)]
    assert A._ew(2, B.evaluations).to_ints() == [x * y % R for x, y in zip(a, b)]
    s = rng.randrange(R)
    assert (A * s).to_ints() == [x * s % R for x in a]
    assert A.to_bytes() == b"".join(pm.be32(x) for x in a)


# ---- the reference's primitive tests, on the GPU ---------------------------------------------------
def test_partial_evaluation_1():  # evaluation_form.rs:315-325
    assert ML([3, 1, 2, 5]).partial_evaluation(F(5), 0) == ML([-2, 21])


def test_partial_evaluation_2():  # evaluation_form.rs:328-358
    p = ML([3, 9, 7, 13, 6, 12, 10, 18])
    assert p.partial_evaluation(2, 0).evaluation([3, 2]) == 57
    assert p.partial_evaluation(3, 1).evaluation([3, 2]) == 72
    assert p.partial_evaluation(1, 2).evaluation([3, 2]) == 38


def test_evaluation():  # evaluation_form.rs:362-405
    assert ML([3, 1, 2, 5]).evaluation([5, 6]) == 136
    assert ML([3, 9, 7, 13, 6, 12, 10, 18]).evaluation([2, 3, 1]) == 39
    assert ML([0, 0, 0, 3, 0, 0, 2, 5]).evaluation([2, 3, 4]) == 48


def test_half_sums_and_hypercube_sum():  # evaluation_form.rs:408-462
    assert ML([0, 0, 0, 2, 2, 2, 2, 4]).split_poly_into_two_and_sum_each_part() == ML([2, 10])
    assert ML([0, 0, 2, 7, 3, 3, 6, 11]).split_poly_into_two_and_sum_each_part() == ML([9, 23])
    assert ML([1, 2, 3, 4, 5, 6, 7, 8]).sum_over_the_boolean_hypercube() == 36


def test_add_mul_distinct():  # evaluation_form.rs:264-312
    a, b = ML([0, 0, 2, 2]), ML([0, 3, 0, 3])
    assert a.add_distinct(b) == ML([0, 3, 0, 3, 0, 3, 0, 3, 2, 5, 2, 5, 2, 5, 2, 5])
    assert a.mul_distinct(b) == ML([0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 0, 6, 0, 6, 0, 6])


def test_element_wise_product():  # composed_multilinear.rs:159-170
    assert zk.ComposedMultilinear([ML([0, 1, 2, 3]), ML([0, 0, 0, 1])]).element_wise_product() == [0, 0, 0, 3]


def test_sum_calculations():  # sumcheck.rs:108-123, composed_sumcheck.rs:108-140, multi_composed_sumcheck.rs:195-214
    s = zk.Sumcheck(ML([0, 0, 0, 2, 2, 2, 2, 4]))
    s.poly_sum()
    assert s.sum == 12
    CS, CM = zk.ComposedSumcheck, zk.ComposedMultilinear
    assert CS.calculate_poly_sum(CM([ML([0, 1, 2, 3]), ML([0, 0, 0, 1])])) == 3
    assert CS.calculate_poly_sum(CM([ML([3, 3, 5, 5]), ML([0, 0, 0, 1])])) == 5
    assert CS.calculate_poly_sum(CM([ML([0, 1, 2, 3])])) == 6
    MP = zk.MultiComposedSumcheckProver
    assert MP.calculate_poly_sum([CM([ML([0, 1, 2, 3])]), CM([ML([0, 0, 0, 1])])]) == 7
    assert MP.calculate_poly_sum([CM([ML([0, 0, 0, 2])]), CM([ML([0, 3, 0, 3])])]) == 8


def test_general_variable_index_fold_vs_oracle():
    rng = random.Random(22)
    ev = [rng.randrange(R) for _ in range(64)]
    for k in range(6):
        r = rng.randrange(R)
        assert zk.Multilinear(ev).partial_evaluation(r, k).to_ints() == cref.partial_evaluation(ev, r, k)
    pts = [rng.randrange(R) for _ in range(6)]
    assert zk.Multilinear(ev).evaluation(pts) == cref.evaluation(ev, pts)


def test_shape_errors_mirror_reference_asserts():
    with pytest.raises(zk.ZkscError) as e:
        zk.Multilinear([1, 2, 3])                       # evaluation_form.rs:16-20
    assert e.value.code == -3
    with pytest.raises(zk.ZkscError):
        ML([1, 2, 3, 4]).evaluation([1])                # :163-167
    with pytest.raises(zk.ZkscError):
        zk.ComposedMultilinear([ML([1, 2]), ML([1, 2, 3, 4])])   # composed_multilinear.rs:15
    with pytest.raises(zk.ZkscError):
        ML([1, 2, 3, 4]).partial_evaluation(1, 2)       # polynomial/src/utils.rs:30-34


# ---- the reference's protocol tests (round trip) + byte parity with the oracle ----------------------
REF_INPUTS = [[0, 0, 2, 7, 3, 3, 6, 11], [0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0], [1, 3, 5, 7, 2, 4, 6, 8, 3, 5, 7, 9, 4, 6, 8, 10]]


@pytest.mark.parametrize("ev", REF_INPUTS)
def test_sumcheck_proofs(ev):  # sumcheck.rs:126-202
    sc = zk.Sumcheck(ML(ev))
    sc.poly_sum()
    proof, ch = sc.prove()
    assert sc.verify(proof)
    o = pm.Sumcheck(pm.Multilinear(ev)); o.poly_sum()
    op, och = o.prove()
    assert sc.sum == o.sum and ch == och
    assert proof.to_bytes() == b"".join(u.to_bytes() for u in op.univariate_poly)
    proof.sum = (proof.sum + 1) % R
    assert not sc.verify(proof)


@pytest.mark.parametrize("ev", REF_INPUTS)
def test_composed_sumcheck_proofs(ev):  # composed_sumcheck.rs:167-241
    cs = zk.ComposedSumcheck(zk.ComposedMultilinear([ML(ev)]))
    proof, ch = cs.prove()
    assert cs.verify(proof, zk.ComposedSumcheck.calculate_poly_sum(proof.poly))
    op, och = pm.ComposedSumcheck(pm.ComposedMultilinear([pm.Multilinear(ev)])).prove()
    assert ch == och and proof.round_polys == op.round_polys


def test_composed_sumcheck_product():  # composed_sumcheck.rs:143-164
    cs = zk.ComposedSumcheck(zk.ComposedMultilinear([ML([3, 3, 5, 5]), ML([0, 0, 0, 1])]))
    proof, _ = cs.prove()
    assert cs.verify(proof, zk.ComposedSumcheck.calculate_poly_sum(proof.poly))
    assert not cs.verify(proof, 6)


def _multi_cases(MLc, CM):
    p1, p2 = MLc([0, 0, 0, 2]), MLc([0, 3, 0, 3])
    yield [CM([p1]), CM([p2])]
    yield [CM([p1]), CM([p2]), CM([p2])]
    yield [CM([p1, p2]), CM([p2, p1])]
    add_i, mul_i = MLc([4, 4, 7, 7, 4, 4, 7, 9]), MLc([3, 3, 3, 4, 3, 3, 5, 6])
    w_b, w_c = MLc([0, 4]), MLc([0, 3])
    yield [CM([add_i.partial_evaluation(2, 0), w_b.add_distinct(w_c)]), CM([mul_i.partial_evaluation(2, 0), w_b.mul_distinct(w_c)])]


@pytest.mark.parametrize("idx", range(4))
def test_multi_composed_sumcheck_proofs(idx):  # multi_composed_sumcheck.rs:217-311
    polys = list(_multi_cases(ML, zk.ComposedMultilinear))[idx]
    opolys = list(_multi_cases(lambda v: pm.Multilinear(v), pm.ComposedMultilinear))[idx]
    s = zk.MultiComposedSumcheckProver.calculate_poly_sum(polys)
    assert s == pm.MultiComposedSumcheckProver.calculate_poly_sum(opolys)
    proof, ch = zk.MultiComposedSumcheckProver.prove(polys, s)
    assert zk.MultiComposedSumcheckVerifier.verify(polys, proof)
    op, och = pm.MultiComposedSumcheckProver.prove(opolys, s)
    assert ch == och and proof.to_bytes() == op.to_bytes()
    assert [rp.monomial for rp in proof.round_polys] == [rp.monomial for rp in op.round_polys]
    pp, pch = zk.MultiComposedSumcheckProver.prove_partial(polys, s)
    opp, opch = pm.MultiComposedSumcheckProver.prove_partial(opolys, s)
    assert pch == opch and pp.to_bytes() == opp.to_bytes()
    sub = zk.MultiComposedSumcheckVerifier.verify_partial(pp)
    assert sub.challenges == pch and sub.sum == sum(p.evaluation(pch) for p in opolys) % R
    if idx == 3:  # SURVEY 8c derived check values
        assert s == 213 and proof.round_polys[0].monomial == [(12, 0), (89, 1), (100, 2)]
        assert ch[0] == 0x716D56C92CA17EBB13DC4AEF7135E1BCDB9FA3005A89F37E8B9A2D9C9BB061E4
    pp.sum = (pp.sum + 1) % R
    with pytest.raises(zk.ZkscError) as e:
        zk.MultiComposedSumcheckVerifier.verify_partial(pp)
    assert e.value.code == -7


# ---- randomized parity sweep against the C oracle ---------------------------------------------------
def _gen(kind, n, rng):
    if kind == "rand":
        return [rng.randrange(R) for _ in range(1 << n)]
    if kind == "small":
        return [rng.randrange(256) for _ in range(1 << n)]
    if kind == "zero":
        return [0] * (1 << n)
    if kind == "ones":
        return [1] * (1 << n)
    if kind == "ramp":
        return [i % 256 for i in range(1 << n)]
    if kind == "max":
        return [R - 1 - (i % 3) for i in range(1 << n)]
    return [rng.randrange(2) for _ in range(1 << n)]


def _prove_gpu(ctx, proto, n, degs, tabs_canon, s):
    mont = np.empty_like(tabs_canon)
    zk.lib().zksc_fr_from_canonical_batch(_lib.p64(tabs_canon), tabs_canon.shape[0], _lib.p64(mont))
    N = 1 << n
    t = zk.Tables.upload(ctx, n, degs, [mont[i * N:(i + 1) * N] for i in range(sum(degs))])
    try:
        assert zk.from_mont(t.poly_sum()[0]) == cref.poly_sum(n, degs, tabs_canon)
        msgs, lens, chal = t.prove(proto, zk.to_mont(s))
        return _lib.proof_to_bytes(proto, msgs[0], lens[0]), (zk.from_mont(chal[0]) if n else [])
    finally:
        t.free()


@pytest.mark.parametrize("seed", range(48))
def test_prove_parity_sweep(ctx, seed):
    rng = random.Random(3000 + seed)
    proto = seed % 4
    n = rng.randint(1, 11)
    degs = [rng.randint(1, 5) for _ in range(rng.randint(1, 3))]
    if proto == 0:
        degs = [1]
    if proto == 1:
        degs = degs[:1]
    kind = ["rand", "small", "zero", "bits", "ones", "ramp", "max"][seed % 7]
    flat = cref.ints_to_canon([v for d in degs for _ in range(d) for v in _gen(kind, n, rng)])
    s = cref.poly_sum(n, degs, flat)
    if seed % 5 == 0:
        s = (s + 7) % R
    assert _prove_gpu(ctx, proto, n, degs, flat, s) == cref.prove(proto, n, degs, flat, s)


@pytest.mark.parametrize("n,degs", [(14, [1]), (15, [2]), (14, [3]), (13, [2, 2]), (12, [4]), (11, [6]), (10, [8]), (12, [2, 3, 1]), (16, [2]), (17, [3])])
def test_prove_partial_parity_larger(ctx, n, degs):
    """sizes where several blocks / the grid-stride loop / lazy accumulation over many pairs are exercised"""
    cref.set_threads(8)
    tabs = np.concatenate([cref.synth_table(99 + n, k, n) for k in range(sum(degs))])
    s = cref.poly_sum(n, degs, tabs)
    t = zk.Tables.synth(ctx, n, degs, 99 + n)
    try:
        assert zk.from_mont(t.poly_sum()[0]) == s
        msgs, lens, chal = t.prove(zk.PROTO_MULTI_PARTIAL, zk.to_mont(s))
        got = _lib.proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs[0], lens[0]), zk.from_mont(chal[0])
    finally:
        t.free()
    assert got == cref.prove(2, n, degs, tabs, s)
    cref.set_threads(1)


def test_low_level_round_api_vs_model(ctx):
    """zksc_round_evals / zksc_bind / zksc_residual against the model, incl. unfused binds"""
    rng = random.Random(31)
    n, degs = 6, [2, 3]
    tabs = [[rng.randrange(R) for _ in range(1 << n)] for _ in range(sum(degs))]
    polys = [pm.ComposedMultilinear([pm.Multilinear(t) for t in tabs[:2]]), pm.ComposedMultilinear([pm.Multilinear(t) for t in tabs[2:]])]
    t = zk.Tables.upload(ctx, n, degs, [zk.to_mont(x) for x in tabs])
    cur = polys
    for rnd in range(n):
        want = [v for p in cur for v in pm.round_evals(p)]
        assert zk.from_mont(t.round_evals()[0]) == want
        r = rng.randrange(R)
        t.bind(zk.to_mont(r))
        cur = [p.partial_evaluation(r, 0) for p in cur]
        if rnd % 2 == 1 and rnd < n - 1:   # two binds in a row: the first is flushed by the stand-alone fold kernel
            r2 = rng.randrange(R)
            t.bind(zk.to_mont(r2))
            cur = [p.partial_evaluation(r2, 0) for p in cur]
            if t.vars_left() == 0:
                break
        if t.vars_left() == 0:
            break
    res = t.residual()[0]
    assert [zk.from_mont(res[k])[0] for k in range(5)] == [m.evaluations[0] for p in cur for m in p.polys]
    t.reset()
    assert zk.from_mont(t.round_evals()[0]) == [v for p in polys for v in pm.round_evals(p)]
    assert t.to_bytes(0) == pm.composed_poly_to_bytes(polys)
    t.free()


def _model_polys(tabs, degs):
    out, k = [], 0
    for d in degs:
        out.append(pm.ComposedMultilinear([pm.Multilinear(t) for t in tabs[k:k + d]]))
        k += d
    return out


def test_resident_tail_kernel_call_patterns(ctx):
    """The persistent tail kernel (K4) is started by round_evals once the tables are small and fed by bind; any other
    call pattern must stop it without losing state: repeated round_evals, binds in a row, residual / reset / a second
    handle / a stand-alone Multilinear operation in the middle of a tail."""
    rng = random.Random(77)
    n, degs = 7, [2, 1, 3]
    tabs = [[rng.randrange(R) for _ in range(1 << n)] for _ in range(sum(degs))]
    t = zk.Tables.upload(ctx, n, degs, [zk.to_mont(x) for x in tabs])
    other = zk.Tables.synth(ctx, 5, [2], 9)
    cur = _model_polys(tabs, degs)

    def model_evals():
        return [v for p in cur for v in pm.round_evals(p)]

    def bind(r):
        nonlocal cur
        t.bind(zk.to_mont(r))
        cur = [p.partial_evaluation(r, 0) for p in cur]

    assert zk.from_mont(t.round_evals()[0]) == model_evals()          # round 0: ordinary launch
    bind(rng.randrange(R))
    assert zk.from_mont(t.round_evals()[0]) == model_evals()          # round 1: the tail kernel starts here
    assert zk.from_mont(t.round_evals()[0]) == model_evals()          # asked again without a bind: tail stopped, recomputed
    bind(rng.randrange(R))
    assert zk.from_mont(t.round_evals()[0]) == model_evals()          # round 2 (tail again)
    bind(rng.randrange(R))
    bind(rng.randrange(R))                                            # two binds in a row with a tail resident
    assert zk.from_mont(t.round_evals()[0]) == model_evals()          # round 4
    bind(rng.randrange(R))
    o = other.poly_sum()                                              # another handle takes the stream mid-tail
    assert zk.from_mont(o[0]) == sum(a * b for a, b in zip(*[cref.canon_to_ints(cref.synth_table(9, k, 5)) for k in range(2)])) % R
    assert zk.from_mont(t.round_evals()[0]) == model_evals()          # round 5
    bind(rng.randrange(R))
    m = zk.Multilinear([1, 2, 3, 4]).partial_evaluation(5, 0)            # a stand-alone operation on the same context mid-tail
    assert m.to_ints() == [(1 + 5 * 2) % R, (2 + 5 * 2) % R]
    res = t.residual()[0]                                             # residual with a challenge posted to the tail
    assert [zk.from_mont(res[k]) for k in range(sum(degs))] == [m.evaluations for p in cur for m in p.polys]
    assert zk.from_mont(t.round_evals()[0]) == model_evals()          # last round, after the residual read
    t.reset()                                                         # and from the top again, as a whole proof
    s = t.poly_sum()
    msgs, lens, chal = t.prove(zk.PROTO_MULTI_PARTIAL, s)
    proof, ch = pm.MultiComposedSumcheckProver.prove_partial(_model_polys(tabs, degs), zk.from_mont(s[0]))
    assert _lib.proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs[0], lens[0]) == proof.to_bytes()
    assert zk.from_mont(chal[0]) == ch
    t.free()
    other.free()


@pytest.mark.parametrize("n,degs,B", [(12, [2], 1), (9, [3, 1], 3), (14, [2, 2], 1)])
def test_proof_after_a_proof_left_half_way(ctx, n, degs, B):
    """A proof abandoned in the middle of the resident kernel's rounds (reset), then a whole proof whose round 0 runs inside the prover: the
    resident kernel is queued behind that round's launch before any challenge is posted, so it must not read the "leave" mark the
    abandoned session put into the mailbox (found by tools/stress.py)."""
    rng = random.Random(5)
    t = zk.Tables.synth(ctx, n, degs, 4242, n_proofs=B)
    s = t.poly_sum()
    first = t.prove(zk.PROTO_MULTI_PARTIAL, s)
    for stop_after in (2, n - 1, 1):
        t.reset()
        for _ in range(stop_after):
            t.round_evals()
            t.bind(zk.to_mont([rng.randrange(R) for _ in range(B)]))
        t.reset()
        again = t.prove(zk.PROTO_MULTI_PARTIAL, s)          # no cached round 0 this time: it is launched inside the prover
        for x, y in zip(first, again):
            assert np.array_equal(x, y)
    t.free()


def test_tail_kernel_off_gives_identical_proofs():
    """ZKSC_NO_TAIL=1 (every round its own launch) and the default (resident tail kernel) must agree byte for byte."""
    import os
    outs = []
    for flag in ("0", "1"):
        os.environ["ZKSC_NO_TAIL"] = flag
        c = zk.Context(0)
        t = zk.Tables.synth(c, 13, [3, 2], 4242, n_proofs=3)
        s = t.poly_sum()
        msgs, lens, chal = t.prove(zk.PROTO_MULTI_PARTIAL, s)
        outs.append((msgs.copy(), lens.copy(), chal.copy()))
        t.free()
        c.close()
    os.environ.pop("ZKSC_NO_TAIL")
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_batched_independent_proofs(ctx):
    """config 5 shape at small size: many independent proofs in one launch per round (incl. > 64 proofs; from 16 proofs on the
    per-proof transcript work of a round runs on the library's host worker pool -- every proof is compared with the oracle)"""
    n, degs = 8, [2]
    for B in (3, 70):
        t = zk.Tables.synth(ctx, n, degs, 500, n_proofs=B)
        sums = t.poly_sum()
        msgs, lens, chal = t.prove(zk.PROTO_MULTI_PARTIAL, sums)
        t.reset()
        again = t.prove(zk.PROTO_MULTI_PARTIAL, sums)
        assert np.array_equal(again[0], msgs) and np.array_equal(again[2], chal)
        for b in range(B):
            tabs = np.concatenate([cref.synth_table(500 + b, k, n) for k in range(2)])
            s = cref.poly_sum(n, degs, tabs)
            assert zk.from_mont(sums[b]) == s
            assert (_lib.proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs[b], lens[b]), zk.from_mont(chal[b])) == cref.prove(2, n, degs, tabs, s)
        t.free()


def test_verifier_oracle_check_on_device(ctx):
    """MultiComposedSumcheckVerifier::verify's oracle check (multi_composed_sumcheck.rs:135-141) = n folds on the GPU"""
    rng = random.Random(33)
    n, degs = 9, [2, 1]
    t = zk.Tables.synth(ctx, n, degs, 41)
    pts = [rng.randrange(R) for _ in range(n)]
    got = zk.from_mont(t.evaluate(zk.to_mont(pts))[0])
    tabs = [cref.canon_to_ints(cref.synth_table(41, k, n)) for k in range(3)]
    want = (cref.evaluation(tabs[0], pts) * cref.evaluation(tabs[1], pts) + cref.evaluation(tabs[2], pts)) % R
    assert got == want
    t.free()


# ---- full-size, size-independent properties (BASELINE.json configs 1 and 2) -------------------------
@pytest.mark.parametrize("n,degs,proto", [(20, [1], zk.PROTO_SUMCHECK), (24, [2], zk.PROTO_MULTI_PARTIAL), (22, [3], zk.PROTO_MULTI_PARTIAL),
                                          (26, [3], zk.PROTO_MULTI_PARTIAL), (20, [2, 2], zk.PROTO_MULTI_PARTIAL)])
def test_full_size_prove_verify_roundtrip(ctx, n, degs, proto):
    """prove -> transcript replay (p(0)+p(1) chain) -> oracle check by folding the tables at the challenges:
    the three must close, and a second prove on the same resident tables must give identical bytes."""
    t = zk.Tables.synth(ctx, n, degs, 12345)
    try:
        s = t.poly_sum()
        msgs, lens, chal = t.prove(proto, s)
        sub, ch2 = _lib.verify_rounds(proto, s[0], msgs[0], lens[0])
        assert np.array_equal(ch2, chal[0])
        assert np.array_equal(t.evaluate(chal)[0], sub)
        t.reset()
        msgs2, lens2, chal2 = t.prove(proto, s)
        assert np.array_equal(msgs, msgs2) and np.array_equal(chal, chal2)
        # spot-check the seeded generator against the host formula at a few indices
        out = np.zeros(4, dtype=np.uint64)
        zk.lib().zksc_synth_entry(12345, 0, 0, _lib.p64(out))
    finally:
        t.free()


def test_double_buffered_refill(ctx):
    """zksc_tables_reupload_begin / _end: refill one handle on the copy stream while another is being proved; the refilled
    handle then proves to the same bytes as a freshly uploaded one, also when _end is left to the next call on the handle."""
    n, degs = 14, [2, 1]
    seeds = [11, 12, 13]
    host = []
    for sd in seeds:
        t = zk.Tables.synth(ctx, n, degs, sd)
        host.append(t.read_local().reshape(sum(degs), 1 << n, 4).copy())
        t.free()
    want = []
    for h in host:
        t = zk.Tables.upload(ctx, n, degs, [h[i] for i in range(h.shape[0])])
        s = t.poly_sum()
        want.append((s.copy(), [a.copy() for a in t.prove(zk.PROTO_MULTI_PARTIAL, s)]))
        t.free()
    pair = [zk.Tables.alloc(ctx, n, degs), zk.Tables.alloc(ctx, n, degs)]
    pair[0].reupload_begin([host[0][i] for i in range(3)])
    for k in range(3):
        cur = pair[k % 2]
        if k != 1:
            cur.reupload_end()            # k == 1: left pending on purpose, poly_sum must finish the refill itself
        if k + 1 < 3:
            pair[(k + 1) % 2].reupload_begin([host[k + 1][i] for i in range(3)])
        s = cur.poly_sum()
        got = cur.prove(zk.PROTO_MULTI_PARTIAL, s)
        assert np.array_equal(s, want[k][0])
        for a, b in zip(got, want[k][1]):
            assert np.array_equal(a, b)
    with pytest.raises(zk.ZkscError):
        pair[0].reupload_begin([host[0][i] for i in range(3)])
        pair[0].reupload_begin([host[0][i] for i in range(3)])   # second begin without end: STATE error
    pair[0].reupload_end()
    for t in pair:
        t.free()


def test_batch_of_64_proofs_roundtrip(ctx):
    """BASELINE config 5 shape (64 independent proofs, batched launches, host worker pool) at 2^16 entries each: every proof's
    transcript replays, its oracle check closes on the device, and proof 0 / 63 equal the oracle's bytes."""
    n, degs, B = 16, [2], 64
    t = zk.Tables.synth(ctx, n, degs, 900, n_proofs=B)
    try:
        sums = t.poly_sum()
        msgs, lens, chal = t.prove(zk.PROTO_MULTI_PARTIAL, sums)
        vals = t.evaluate(chal)
        for b in range(B):
            sub, ch2 = _lib.verify_rounds(zk.PROTO_MULTI_PARTIAL, sums[b], msgs[b], lens[b])
            assert np.array_equal(ch2, chal[b]) and np.array_equal(vals[b], sub)
        for b in (0, B - 1):
            tabs = np.concatenate([cref.synth_table(900 + b, k, n) for k in range(2)])
            s = cref.poly_sum(n, degs, tabs)
            assert (_lib.proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs[b], lens[b]), zk.from_mont(chal[b])) == cref.prove(2, n, degs, tabs, s)
    finally:
        t.free()


def test_sumcheck_utils_helpers():  # sumcheck/src/utils.rs
    """the byte / hypercube helpers of the reference's sumcheck::utils, device-backed, against the Python model"""
    from zk_cryptography_b200 import utils as U
    rng = random.Random(5)
    assert U.convert_field_to_byte(1) == bytes(31) + b"\x01" and U.convert_field_to_byte(100)[-1] == 100      # utils.rs tests
    assert U.boolean_hypercube(2) == [[0, 0], [0, 1], [1, 0], [1, 1]]
    ev = [rng.randrange(R) for _ in range(16)]
    f = ML(ev)
    got = U.skip_first_and_sum_all(f)
    assert got.to_ints() == [sum(ev[:8]) % R, sum(ev[8:]) % R] == f.split_poly_into_two_and_sum_each_part().to_ints()
    assert U.skip_first_and_sum_all(ML([0, 0, 2, 7, 3, 3, 6, 11])).to_ints() == [9, 23]      # evaluation_form.rs:408-438 values
    a, b = [rng.randrange(R) for _ in range(8)], [rng.randrange(R) for _ in range(8)]
    polys = [zk.ComposedMultilinear([ML(a), ML(b)]), zk.ComposedMultilinear([ML(b)])]
    opolys = [pm.ComposedMultilinear([pm.Multilinear(a), pm.Multilinear(b)]), pm.ComposedMultilinear([pm.Multilinear(b)])]
    assert U.sum_over_boolean_hypercube(polys) == pm.sum_over_boolean_hypercube(opolys)
    assert U.composed_poly_to_bytes(polys) == pm.composed_poly_to_bytes(opolys)
    assert U.vec_to_bytes(a) == pm.vec_to_bytes(a)
    assert U.convert_round_poly_to_uni_poly_format([5, 6, 7]) == [(0, 5), (1, 6), (2, 7)]
