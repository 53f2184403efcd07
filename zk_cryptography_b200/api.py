"""Host-side mirror of the reference's public API for the sumcheck path, over the zksc C ABI.

Names, argument meaning and error behaviour follow the reference (paths relative to its root):
  Multilinear                      polynomial/src/multilinear/evaluation_form.rs
  ComposedMultilinear              polynomial/src/composed/composed_multilinear.rs
  SparseUnivariatePolynomial       polynomial/src/univariate/sparse_univariate.rs   (host, tiny data)
  FiatShamirTranscript             transcripts/fiat-shamir/src/fiat_shamir.rs       (host)
  Sumcheck                         sumcheck/src/sumcheck.rs
  ComposedSumcheck                 sumcheck/src/composed/composed_sumcheck.rs
  MultiComposedSumcheckProver/Verifier   sumcheck/src/composed/multi_composed_sumcheck.rs
Where the reference panics (`assert!`) this raises ZkscError(SHAPE); where it returns
Err("Verification failed") this raises ZkscError(VERIFY).  All table-sized arithmetic runs in CUDA.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (PROTO_COMPOSED, PROTO_MULTI_FULL, PROTO_MULTI_PARTIAL, PROTO_SUMCHECK, Context, Tables, ZkscError, from_mont, lib, p8,
                   p64, to_mont)

_DEFAULT_CTX = None


def default_context():
    global _DEFAULT_CTX
    if _DEFAULT_CTX is None:
        _DEFAULT_CTX = Context(0)
    return _DEFAULT_CTX


def set_default_context(ctx):
    global _DEFAULT_CTX
    _DEFAULT_CTX = ctx


def _as_table(evaluations):
    if isinstance(evaluations, np.ndarray) and evaluations.dtype == np.uint64 and evaluations.ndim == 2 and evaluations.shape[1] == 4:
        return np.ascontiguousarray(evaluations)
    return to_mont([int(v) for v in evaluations])


def _scalar(x):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint64).reshape(4)
    return to_mont(int(x))


class Multilinear:
    """Dense evaluation table over {0,1}^n; variable 0 is the most significant index bit."""

    def __init__(self, evaluations):  # evaluation_form.rs:12-26
        ev = _as_table(evaluations)
        n = ev.shape[0]
        if n == 0 or n & (n - 1):
            raise ZkscError(-3, "Number of evaluations must be a power of 2")
        self.n_vars = n.bit_length() - 1
        self.evaluations = ev

    new = classmethod(lambda cls, ev: cls(ev))

    def __eq__(self, o):
        return self.n_vars == o.n_vars and np.array_equal(self.evaluations, o.evaluations)

    def to_ints(self):
        return from_mont(self.evaluations)

    def _ctx(self):
        return default_context()

    def partial_evaluation(self, eval_point, variable_index):  # :123-141
        ctx = self._ctx()
        n = self.evaluations.shape[0]
        out = np.zeros((max(n // 2, 1), 4), dtype=np.uint64)
        r = _scalar(eval_point)
        ctx.check(lib().zksc_ml_partial_evaluation(ctx._h, p64(self.evaluations), n, p64(r), variable_index, p64(out)))
        return Multilinear(out)

    def partial_evaluations(self, points, variable_indices):  # :143-159
        if len(points) != len(variable_indices):
            raise ZkscError(-3, "The length of evaluation_points and variable_indices should be the same")
        e = self
        for p, k in zip(points, variable_indices):
            e = e.partial_evaluation(p, k)
        return e

    def evaluation(self, evaluation_points):  # :162-175
        if len(evaluation_points) != self.n_vars:
            raise ZkscError(-3, "Number of evaluation points must match the number of variables")
        ctx = self._ctx()
        pts = np.stack([_scalar(p) for p in evaluation_points]) if evaluation_points else np.zeros((0, 4), dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        ctx.check(lib().zksc_ml_evaluation(ctx._h, p64(self.evaluations), self.evaluations.shape[0], p64(pts), len(evaluation_points), p64(out)))
        return from_mont(out)

    def _outer(self, rhs, mul):
        ctx = self._ctx()
        na, nb = self.evaluations.shape[0], rhs.evaluations.shape[0]
        out = np.zeros((na * nb, 4), dtype=np.uint64)
        ctx.check(lib().zksc_ml_outer(ctx._h, mul, p64(self.evaluations), na, p64(rhs.evaluations), nb, p64(out)))
        return Multilinear(out)

    def add_distinct(self, rhs):  # :28-39
        return self._outer(rhs, 0)

    def mul_distinct(self, rhs):  # :41-52
        return self._outer(rhs, 1)

    def _ew(self, op, other):
        ctx = self._ctx()
        n = self.evaluations.shape[0]
        out = np.zeros((n, 4), dtype=np.uint64)
        ctx.check(lib().zksc_ml_elementwise(ctx._h, op, p64(self.evaluations), p64(other), n, p64(out)))
        return Multilinear(out)

    def __add__(self, rhs):  # impl Add :178-194
        return self._ew(0, rhs.evaluations)

    def __sub__(self, rhs):  # impl Sub :209-225
        return self._ew(1, rhs.evaluations)

    def __mul__(self, scalar):  # impl Mul<F> :235-251
        return self._ew(3, _scalar(scalar).reshape(1, 4))

    def to_bytes(self):  # :54-62
        t = Tables.upload(self._ctx(), self.n_vars, [1], [self.evaluations])
        try:
            return t.to_bytes(0)
        finally:
            t.free()

    def split_poly_into_two_and_sum_each_part(self):  # :68-74
        if self.n_vars == 0:
            raise ZkscError(-3, "a constant has no halves")
        t = Tables.upload(self._ctx(), self.n_vars, [1], [self.evaluations])
        try:
            return Multilinear(t.round_evals()[0])
        finally:
            t.free()

    def sum_over_the_boolean_hypercube(self):  # :80-84
        t = Tables.upload(self._ctx(), self.n_vars, [1], [self.evaluations])
        try:
            return from_mont(t.poly_sum()[0])
        finally:
            t.free()


class ComposedMultilinear:
    """Product of same-arity multilinear tables."""

    def __init__(self, polys):  # composed_multilinear.rs:13-18
        if not polys or any(p.n_vars != polys[0].n_vars for p in polys):
            raise ZkscError(-3, "all factors must have the same number of variables")
        self.polys = list(polys)

    new = classmethod(lambda cls, polys: cls(polys))

    def n_vars(self):
        return self.polys[0].n_vars

    def max_degree(self):  # :101-103
        return len(self.polys)

    def to_bytes(self):  # :40-48
        return b"".join(p.to_bytes() for p in self.polys)

    def partial_evaluation(self, point, variable_index):  # :63-75
        return ComposedMultilinear([p.partial_evaluation(point, variable_index) for p in self.polys])

    def evaluation(self, points):  # :52-61
        res = 1
        for p in self.polys:
            res = res * p.evaluation(points) % _lib.R_MOD
        return res

    def element_wise_product(self):  # :105-111
        acc = self.polys[0]
        for p in self.polys[1:]:
            acc = acc._ew(2, p.evaluations)
        return acc.to_ints()

    def element_wise_add(self):  # :113-119
        acc = self.polys[0]
        for p in self.polys[1:]:
            acc = acc + p
        return acc.to_ints()


def _upload(polys):
    """list[ComposedMultilinear] -> Tables (one proof)."""
    n = polys[0].n_vars()
    if any(p.n_vars() != n for p in polys):
        raise ZkscError(-3, "all products must have the same number of variables")
    degs = [p.max_degree() for p in polys]
    tabs = [m.evaluations for p in polys for m in p.polys]
    return Tables.upload(default_context(), n, degs, tabs)


class FiatShamirTranscript:  # fiat_shamir.rs:5-40
    def __init__(self):
        self._h = lib().zksc_transcript_new()

    def commit(self, data):
        data = bytes(data)
        lib().zksc_transcript_commit(self._h, data, len(data))

    def challenge(self):
        out = np.zeros(32, dtype=np.uint8)
        lib().zksc_transcript_challenge(self._h, p8(out))
        return out.tobytes()

    def evaluate_challenge_into_field(self):
        out = np.zeros(4, dtype=np.uint64)
        lib().zksc_transcript_challenge_field(self._h, p64(out))
        return from_mont(out)

    def evaluate_n_challenge_into_field(self, n):
        return [self.evaluate_challenge_into_field() for _ in range(n)]

    def __del__(self):
        try:
            lib().zksc_transcript_free(self._h)
        except Exception:
            pass


class SparseUnivariatePolynomial:  # sparse_univariate.rs
    def __init__(self, monomial=None):
        self.monomial = list(monomial or [])  # [(coeff, pow)] canonical ints

    @staticmethod
    def zero():
        return SparseUnivariatePolynomial([])

    def __eq__(self, o):
        return self.monomial == o.monomial

    def _arr(self):
        flat = [v for cp in self.monomial for v in cp]
        return to_mont(flat) if flat else np.zeros((0, 4), dtype=np.uint64)

    @staticmethod
    def _from_arr(arr, n):
        v = from_mont(arr[:2 * n]) if n else []
        return SparseUnivariatePolynomial([(v[2 * i], v[2 * i + 1]) for i in range(n)])

    @staticmethod
    def interpolation_evals(ys):
        """interpolation over x = 0..len(ys)-1 (the only form the sumcheck path uses, :40-63)."""
        y = to_mont([int(v) for v in ys])
        out = np.zeros((2 * len(ys), 4), dtype=np.uint64)
        n = lib().zksc_sparse_interpolate(p64(y), len(ys), p64(out))
        return SparseUnivariatePolynomial._from_arr(out, n)

    def __add__(self, rhs):  # :159-203
        a, b = self._arr(), rhs._arr()
        out = np.zeros((2 * (len(self.monomial) + len(rhs.monomial)) + 2, 4), dtype=np.uint64)
        n = lib().zksc_sparse_add(p64(a), len(self.monomial), p64(b), len(rhs.monomial), p64(out))
        return SparseUnivariatePolynomial._from_arr(out, n)

    def evaluate(self, point):  # :90-106
        out = np.zeros(4, dtype=np.uint64)
        a = self._arr()
        lib().zksc_sparse_evaluate(p64(a), len(self.monomial), p64(_scalar(point)), p64(out))
        return from_mont(out)

    def to_bytes(self):  # :27-34
        return b"".join((c % _lib.R_MOD).to_bytes(32, "big") + (p % _lib.R_MOD).to_bytes(32, "big") for c, p in self.monomial)


# ---- sumcheck/src/sumcheck.rs ---------------------------------------------------------------------
def _pack_rounds(rows, per=1):
    """Public proof fields -> the (msgs, lens) arrays of the C ABI.  rows: per round, the list of canonical ints the round
    message holds (evaluations, or coeff/pow pairs flattened when per = 2).  The proof objects keep NO hidden copy of the
    prover's output: what a verifier checks (and what to_bytes serialises) is exactly what the public fields say, as in the
    reference, whose verifiers read proof.round_polys / proof.univariate_poly."""
    n = len(rows)
    stride = max([len(r) for r in rows] + [per])
    msgs = np.zeros((n, stride, 4), dtype=np.uint64)
    lens = np.zeros(n, dtype=np.uint32)
    for i, r in enumerate(rows):
        if len(r) % per:
            raise ZkscError(-3, "malformed round message")
        if r:
            msgs[i, :len(r)] = to_mont([int(v) for v in r])
        lens[i] = len(r) // per
    return msgs, lens


class SumcheckProof:  # :11-15
    def __init__(self, poly, s, univariate_poly):
        self.poly, self.sum, self.univariate_poly = poly, s, univariate_poly

    def _rounds(self):
        return _pack_rounds([u.to_ints() for u in self.univariate_poly])

    def to_bytes(self):
        """concatenation of the per-round messages as absorbed by the transcript"""
        return _lib.proof_to_bytes(PROTO_SUMCHECK, *self._rounds())


class Sumcheck:
    def __init__(self, poly):  # :18-23
        self.poly = poly
        self.sum = 0

    new = classmethod(lambda cls, poly: cls(poly))

    def poly_sum(self):  # :25-27
        self.sum = self.poly.sum_over_the_boolean_hypercube()

    def prove(self):  # :29-61
        t = _upload([ComposedMultilinear([self.poly])])
        try:
            msgs, lens, chal = t.prove(PROTO_SUMCHECK, to_mont(self.sum))
        finally:
            t.free()
        unis = [Multilinear(msgs[0, r, :2]) for r in range(self.poly.n_vars)]
        return SumcheckProof(self.poly, self.sum, unis), from_mont(chal[0]) if self.poly.n_vars else []

    def verify(self, proof):  # :63-95
        try:
            sub, chal = _lib.verify_rounds(PROTO_SUMCHECK, to_mont(proof.sum), *proof._rounds())
        except ZkscError as e:
            if e.code == -7:
                return False
            raise
        return proof.poly.evaluation(from_mont(chal) if len(chal) else []) == from_mont(sub)


# ---- sumcheck/src/composed/composed_sumcheck.rs ------------------------------------------------------
class ComposedSumcheckProof:  # :15-18
    def __init__(self, poly, round_polys):
        self.poly, self.round_polys = poly, round_polys

    def _rounds(self):
        return _pack_rounds(self.round_polys)

    def to_bytes(self):
        return _lib.proof_to_bytes(PROTO_COMPOSED, *self._rounds())


class ComposedSumcheck:
    def __init__(self, poly):  # :21-26
        self.poly = poly
        self.sum = 0

    new = classmethod(lambda cls, poly: cls(poly))

    @staticmethod
    def calculate_poly_sum(poly):  # :28-30
        t = _upload([poly])
        try:
            return from_mont(t.poly_sum()[0])
        finally:
            t.free()

    def prove(self):  # :32-67
        t = _upload([self.poly])
        try:
            msgs, lens, chal = t.prove(PROTO_COMPOSED)
        finally:
            t.free()
        n, d = self.poly.n_vars(), self.poly.max_degree()
        rps = [from_mont(msgs[0, r, :d + 1]) for r in range(n)]
        return ComposedSumcheckProof(self.poly, rps), from_mont(chal[0]) if n else []

    def verify(self, proof, s):  # :69-95
        try:
            sub, chal = _lib.verify_rounds(PROTO_COMPOSED, to_mont(s), *proof._rounds())
        except ZkscError as e:
            if e.code == -7:
                return False
            raise
        return proof.poly.evaluation(from_mont(chal) if len(chal) else []) == from_mont(sub)


# ---- sumcheck/src/composed/multi_composed_sumcheck.rs --------------------------------------------------
class MultiComposedProof:
    """ComposedSumcheckProof of multi_composed_sumcheck.rs:12-16 (round_polys: sparse polynomials, sum)."""

    def __init__(self, round_polys, s):
        self.round_polys, self.sum = round_polys, s

    def _rounds(self):
        return _pack_rounds([[v for cp in rp.monomial for v in cp] for rp in self.round_polys], per=2)

    def to_bytes(self):  # :24-32
        return _lib.proof_to_bytes(PROTO_MULTI_PARTIAL, *self._rounds())


class SubClaim:  # :18-22
    def __init__(self, s, challenges):
        self.sum, self.challenges = s, challenges


class MultiComposedSumcheckProver:
    @staticmethod
    def calculate_poly_sum(poly):  # :37-45
        t = _upload(poly)
        try:
            return from_mont(t.poly_sum()[0])
        finally:
            t.free()

    @staticmethod
    def _prove(poly, s, protocol):
        t = _upload(poly)
        try:
            msgs, lens, chal = t.prove(protocol, to_mont(s))
        finally:
            t.free()
        n = poly[0].n_vars()
        rps = []
        for r in range(n):
            k = int(lens[0, r])
            v = from_mont(msgs[0, r, :2 * k]) if k else []
            rps.append(SparseUnivariatePolynomial([(v[2 * i], v[2 * i + 1]) for i in range(k)]))
        return MultiComposedProof(rps, s % _lib.R_MOD), (from_mont(chal[0]) if n else [])

    @staticmethod
    def prove(poly, s):  # :47-54
        return MultiComposedSumcheckProver._prove(poly, s, PROTO_MULTI_FULL)

    @staticmethod
    def prove_partial(poly, s):  # :56-62
        return MultiComposedSumcheckProver._prove(poly, s, PROTO_MULTI_PARTIAL)


class MultiComposedSumcheckVerifier:
    @staticmethod
    def verify(poly, proof):  # :126-142
        prefix = b"".join(p.to_bytes() for p in poly)  # composed_poly_to_bytes
        sub, chal = _lib.verify_rounds(PROTO_MULTI_FULL, to_mont(proof.sum), *proof._rounds(), prefix)
        t = _upload(poly)
        try:
            val = from_mont(t.evaluate(chal)[0]) if len(chal) else from_mont(t.poly_sum()[0])
        finally:
            t.free()
        return val == from_mont(sub)

    @staticmethod
    def verify_partial(proof):  # :143-149
        sub, chal = _lib.verify_rounds(PROTO_MULTI_PARTIAL, to_mont(proof.sum), *proof._rounds())
        return SubClaim(from_mont(sub), from_mont(chal) if len(chal) else [])
