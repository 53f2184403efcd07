"""TEST INFRASTRUCTURE ONLY -- independent Python big-int restatement of the reference's sumcheck path.

This file is the *second* oracle (the first is oracle/zkref.c).  It is written with plain Python
integers and hashlib so that it shares no arithmetic code with either the C oracle or the CUDA
product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import it; the product path (zk_cryptography_b200/) never does.

PARITY STATUS: "parity unpinned" at the protocol level.  The reference (aagbotemi/zk-cryptography,
pure Rust) cannot be built in this image (no cargo/rustc, no vendored crates), its field arithmetic
lives in the third-party crates ark-ff 0.4.2 / ark-test-curves 0.4.2 (bls12_381::Fr) and its hash in
sha2 0.10.8, and none of its tests pins a challenge, a round polynomial or proof bytes.  What IS
pinned: every primitive known-answer test of the reference (fold, evaluate, half sums, element-wise
product, Lagrange interpolation, be32 serialisation, hypercube sums) -- see tests/test_oracle_kat.py.

Every function cites the reference file:line it restates (paths relative to /root/reference).
Field: BLS12-381 scalar field Fr; canonical residues as Python ints in [0, R_MOD).
"""
import hashlib

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def fr(x):
    """F::from(i) for (possibly negative) small integers, ark-ff semantics: value mod r."""
    return x % R_MOD


def be32(x):
    """element.into_bigint().to_bytes_be() -- sumcheck/src/utils.rs:7-9 (32 bytes, big endian, canonical)."""
    return (x % R_MOD).to_bytes(32, "big")


def from_be_bytes_mod_order(b):
    """F::from_be_bytes_mod_order -- used at transcripts/fiat-shamir/src/fiat_shamir.rs:28,35."""
    return int.from_bytes(b, "big") % R_MOD


# --------------------------------------------------------------------------------------------
# transcripts/fiat-shamir/src/fiat_shamir.rs:5-40
# --------------------------------------------------------------------------------------------
class FiatShamirTranscript:
    def __init__(self):  # :11-15
        self.hasher = hashlib.sha256()

    def commit(self, new_data):  # :17-19
        self.hasher.update(bytes(new_data))

    def challenge(self):  # :21-25  finalize_reset, then feed the digest back
        response = self.hasher.digest()
        self.hasher = hashlib.sha256()
        self.hasher.update(response)
        return response

    def evaluate_challenge_into_field(self):  # :27-29
        return from_be_bytes_mod_order(self.challenge())

    def evaluate_n_challenge_into_field(self, n):  # :31-39
        return [from_be_bytes_mod_order(self.challenge()) for _ in range(n)]


# --------------------------------------------------------------------------------------------
# polynomial/src/utils.rs:26-53
# --------------------------------------------------------------------------------------------
def pick_pairs_with_random_index(num_of_evaluations, variable_index):
    assert num_of_evaluations % 2 == 0, "n must be even"
    assert variable_index < num_of_evaluations // 2, "variable_index must be less than n/2"
    result = []
    iters = 1 << variable_index
    for _ in range(iters):
        rnd = []
        half = (num_of_evaluations // iters) // 2
        for y1 in range(half):
            rnd.append((y1 + len(result) * 2, half + y1 + len(result) * 2))
        result.extend(rnd)
    return result


# --------------------------------------------------------------------------------------------
# polynomial/src/multilinear/evaluation_form.rs
# --------------------------------------------------------------------------------------------
class Multilinear:
    def __init__(self, evaluations):  # :12-26
        evaluations = [e % R_MOD for e in evaluations]
        n = len(evaluations)
        n_vars = n.bit_length() - 1 if n > 0 else 0
        assert n > 0 and (1 << n_vars) == n, "Number of evaluations must be a power of 2"
        self.n_vars = n_vars
        self.evaluations = evaluations

    def __eq__(self, o):
        return self.n_vars == o.n_vars and self.evaluations == o.evaluations

    def add_distinct(self, rhs):  # :28-39
        return Multilinear([(a + b) % R_MOD for a in self.evaluations for b in rhs.evaluations])

    def mul_distinct(self, rhs):  # :41-52
        return Multilinear([(a * b) % R_MOD for a in self.evaluations for b in rhs.evaluations])

    def to_bytes(self):  # :54-62
        return b"".join(be32(e) for e in self.evaluations)

    def split_poly_into_two_and_sum_each_part(self):  # :68-74
        mid = len(self.evaluations) // 2
        return Multilinear([sum(self.evaluations[:mid]) % R_MOD, sum(self.evaluations[mid:]) % R_MOD])

    def sum_over_the_boolean_hypercube(self):  # :80-84
        return sum(self.evaluations) % R_MOD

    def partial_evaluation(self, eval_point, variable_index):  # :123-141
        ev = self.evaluations
        res = []
        for (i, j) in pick_pairs_with_random_index(len(ev), variable_index):
            y1, y2 = ev[i], ev[j]
            res.append((eval_point * y2 + (1 - eval_point) * y1) % R_MOD)  # :133
        m = Multilinear.__new__(Multilinear)
        m.n_vars = self.n_vars - 1
        m.evaluations = res
        return m

    def partial_evaluations(self, points, variable_indices):  # :143-159
        if len(points) != len(variable_indices):
            raise ValueError("The length of evaluation_points and variable_indices should be the same")
        e = self
        for p, k in zip(points, variable_indices):
            e = e.partial_evaluation(p, k)
        return e

    def evaluation(self, points):  # :162-175
        assert len(points) == self.n_vars, "Number of evaluation points must match the number of variables"
        e = self
        for p in points:
            e = e.partial_evaluation(p, 0)
        return e.evaluations[0]

    def scalar_mul(self, s):  # impl Mul<F> :235-251
        return Multilinear([(e * s) % R_MOD for e in self.evaluations])

    def add(self, rhs):  # impl Add :178-194
        return Multilinear([(a + b) % R_MOD for a, b in zip(self.evaluations, rhs.evaluations)])

    def sub(self, rhs):  # impl Sub :209-225
        return Multilinear([(a - b) % R_MOD for a, b in zip(self.evaluations, rhs.evaluations)])


# --------------------------------------------------------------------------------------------
# polynomial/src/composed/composed_multilinear.rs
# --------------------------------------------------------------------------------------------
class ComposedMultilinear:
    def __init__(self, polys):  # :13-18
        n = polys[0].n_vars
        assert all(p.n_vars == n for p in polys)
        self.polys = list(polys)

    def n_vars(self):  # :20-22
        return self.polys[0].n_vars

    def to_bytes(self):  # :40-48
        return b"".join(p.to_bytes() for p in self.polys)

    def evaluation(self, points):  # :52-61
        res = 1
        for p in self.polys:
            res = res * p.evaluation(points) % R_MOD
        return res

    def partial_evaluation(self, point, variable_index):  # :63-75
        return ComposedMultilinear([p.partial_evaluation(point, variable_index) for p in self.polys])

    def max_degree(self):  # :101-103
        return len(self.polys)

    def element_wise_product(self):  # :105-111
        n = len(self.polys[0].evaluations)
        out = []
        for i in range(n):
            v = 1
            for p in self.polys:
                v = v * p.evaluations[i] % R_MOD
            out.append(v)
        return out

    def element_wise_add(self):  # :113-119
        n = len(self.polys[0].evaluations)
        return [sum(p.evaluations[i] for p in self.polys) % R_MOD for i in range(n)]


# --------------------------------------------------------------------------------------------
# polynomial/src/utils.rs:78-100 and polynomial/src/univariate/sparse_univariate.rs
# --------------------------------------------------------------------------------------------
def lagrange_basis(points, i):  # utils.rs:78-100
    l_i = [1]
    for j, (x_j, _) in enumerate(points):
        if i != j:
            new = [0] * (len(l_i) + 1)
            for k, c in enumerate(l_i):
                new[k] = (new[k] - c * x_j) % R_MOD
                new[k + 1] = (new[k + 1] + c) % R_MOD
            l_i = new
    denom = 1
    for j, (x_j, _) in enumerate(points):
        if j != i:
            denom = denom * (points[i][0] - x_j) % R_MOD
    inv = pow(denom, R_MOD - 2, R_MOD)
    return [c * inv % R_MOD for c in l_i]


class SparseUnivariatePolynomial:
    """monomial: list of (coeff, pow), both field elements -- sparse_univariate.rs:11-20."""

    def __init__(self, monomial=None):
        self.monomial = list(monomial or [])

    @staticmethod
    def new(data):  # :67-88 (flat [coeff, pow, coeff, pow, ...])
        mono = []
        for n in range(0, len(data), 2):
            if n < len(data) - 1:
                mono.append((data[n] % R_MOD, data[n + 1] % R_MOD))
            else:
                mono.append((data[n] % R_MOD, 0))
        return SparseUnivariatePolynomial(mono)

    @staticmethod
    def zero():  # :23-25
        return SparseUnivariatePolynomial([])

    def __eq__(self, o):
        return self.monomial == o.monomial

    def to_bytes(self):  # :27-34
        return b"".join(be32(c) + be32(p) for c, p in self.monomial)

    @staticmethod
    def interpolation(points):  # :40-63 ; zero coefficients are dropped (:52-60)
        result = [0] * len(points)
        for i, (_, y_i) in enumerate(points):
            l_i = [c * y_i % R_MOD for c in lagrange_basis(points, i)]
            for k, c in enumerate(l_i):
                result[k] = (result[k] + c) % R_MOD
        return SparseUnivariatePolynomial([(c, fr(pw)) for pw, c in enumerate(result) if c != 0])

    def evaluate(self, point):  # :90-106
        acc = 0
        for c, p in self.monomial:
            acc = (acc + c * pow(point, p, R_MOD)) % R_MOD
        return acc

    def add(self, rhs):  # impl Add :159-203 ; ordered merge by pow, zero sums are kept
        out = []
        li, ri = 0, 0
        L, Rr = self.monomial, rhs.monomial
        while li < len(L) or ri < len(Rr):
            if li < len(L) and ri < len(Rr):
                l, r_ = L[li], Rr[ri]
                if l[1] == r_[1]:
                    out.append(((l[0] + r_[0]) % R_MOD, l[1]))
                    li += 1
                    ri += 1
                elif l[1] < r_[1]:
                    out.append(l)
                    li += 1
                else:
                    out.append(r_)
                    ri += 1
            elif li < len(L):
                out.append(L[li])
                li += 1
            else:
                out.append(Rr[ri])
                ri += 1
        return SparseUnivariatePolynomial(out)


# --------------------------------------------------------------------------------------------
# sumcheck/src/utils.rs
# --------------------------------------------------------------------------------------------
def convert_round_poly_to_uni_poly_format(round_poly):  # :29-35
    return [(fr(i), v) for i, v in enumerate(round_poly)]


def vec_to_bytes(poly):  # :37-43
    return b"".join(be32(p) for p in poly)


def composed_poly_to_bytes(polys):  # :53-59
    return b"".join(p.to_bytes() for p in polys)


def sum_over_boolean_hypercube(polys):  # :45-51
    ev = [f.element_wise_product() for f in polys]
    return sum(sum(v[i] for v in ev) for i in range(len(ev[0]))) % R_MOD


# --------------------------------------------------------------------------------------------
# sumcheck/src/sumcheck.rs
# --------------------------------------------------------------------------------------------
class SumcheckProof:
    def __init__(self, poly, s, univariate_poly):
        self.poly, self.sum, self.univariate_poly = poly, s, univariate_poly


class Sumcheck:
    def __init__(self, poly):  # :18-23
        self.poly = poly
        self.sum = 0

    def poly_sum(self):  # :25-27
        self.sum = sum(self.poly.evaluations) % R_MOD

    def prove(self):  # :29-61
        uni_polys = []
        t = FiatShamirTranscript()
        t.commit(be32(self.sum))
        challenges = []
        cur = self.poly
        for _ in range(self.poly.n_vars):
            uni = cur.split_poly_into_two_and_sum_each_part()
            t.commit(uni.to_bytes())
            uni_polys.append(uni)
            r = t.evaluate_challenge_into_field()
            challenges.append(r)
            cur = cur.partial_evaluation(r, 0)
        return SumcheckProof(self.poly, self.sum, uni_polys), challenges

    def verify(self, proof):  # :63-95
        t = FiatShamirTranscript()
        t.commit(be32(proof.sum))
        claimed = proof.sum
        challenges = []
        for i in range(proof.poly.n_vars):
            uni = proof.univariate_poly[i]
            if (uni.evaluation([0]) + uni.evaluation([1])) % R_MOD != claimed:
                return False
            t.commit(uni.to_bytes())
            c = t.evaluate_challenge_into_field()
            challenges.append(c)
            claimed = uni.evaluation([c])
        return proof.poly.evaluation(challenges) == claimed


# --------------------------------------------------------------------------------------------
# sumcheck/src/composed/composed_sumcheck.rs
# --------------------------------------------------------------------------------------------
def round_evals(p):
    """The round-evaluation idiom (composed_sumcheck.rs:41-49, multi_composed_sumcheck.rs:81-89):
    for i in 0..=max_degree: partial_evaluation(F::from(i), 0).element_wise_product().sum()."""
    out = []
    for i in range(p.max_degree() + 1):
        out.append(sum(p.partial_evaluation(fr(i), 0).element_wise_product()) % R_MOD)
    return out


class ComposedSumcheckProofEvals:
    """composed_sumcheck.rs:15-18 (round_polys are evaluation vectors)."""

    def __init__(self, poly, round_polys):
        self.poly, self.round_polys = poly, round_polys


class ComposedSumcheck:
    def __init__(self, poly):  # :21-26
        self.poly = poly
        self.sum = 0

    @staticmethod
    def calculate_poly_sum(poly):  # :28-30
        return sum(poly.element_wise_product()) % R_MOD

    def prove(self):  # :32-67 ; NO sum absorbed
        t = FiatShamirTranscript()
        cur = self.poly
        round_polys, challenges = [], []
        for _ in range(self.poly.n_vars()):
            rp = round_evals(cur)
            t.commit(vec_to_bytes(rp))
            r = t.evaluate_challenge_into_field()
            challenges.append(r)
            round_polys.append(rp)
            cur = cur.partial_evaluation(r, 0)
        return ComposedSumcheckProofEvals(self.poly, round_polys), challenges

    def verify(self, proof, s):  # :69-95
        t = FiatShamirTranscript()
        claimed = s % R_MOD
        challenges = []
        for rp in proof.round_polys:
            t.commit(vec_to_bytes(rp))
            c = t.evaluate_challenge_into_field()
            challenges.append(c)
            uni = SparseUnivariatePolynomial.interpolation(convert_round_poly_to_uni_poly_format(rp))
            if claimed != (uni.evaluate(0) + uni.evaluate(1)) % R_MOD:
                return False
            claimed = uni.evaluate(c)
        return proof.poly.evaluation(challenges) == claimed


# --------------------------------------------------------------------------------------------
# sumcheck/src/composed/multi_composed_sumcheck.rs
# --------------------------------------------------------------------------------------------
class ComposedSumcheckProof:
    """multi_composed_sumcheck.rs:12-16."""

    def __init__(self, round_polys, s):
        self.round_polys, self.sum = round_polys, s

    def to_bytes(self):  # :24-32
        return b"".join(rp.to_bytes() for rp in self.round_polys)


class SubClaim:  # :18-22
    def __init__(self, s, challenges):
        self.sum, self.challenges = s, challenges


class MultiComposedSumcheckProver:
    @staticmethod
    def calculate_poly_sum(polys):  # :37-45
        return sum(ComposedSumcheck.calculate_poly_sum(p) for p in polys) % R_MOD

    @staticmethod
    def prove(polys, s):  # :47-54
        t = FiatShamirTranscript()
        t.commit(composed_poly_to_bytes(polys))
        return MultiComposedSumcheckProver.prove_internal(polys, s, t)

    @staticmethod
    def prove_partial(polys, s):  # :56-62
        t = FiatShamirTranscript()
        return MultiComposedSumcheckProver.prove_internal(polys, s, t)

    @staticmethod
    def prove_internal(polys, s, t):  # :64-120
        t.commit(be32(s))
        cur = list(polys)
        round_polys, challenges = [], []
        for _ in range(polys[0].n_vars()):
            round_poly = SparseUnivariatePolynomial.zero()
            for p in cur:
                vec = round_evals(p)
                rip = SparseUnivariatePolynomial.interpolation(convert_round_poly_to_uni_poly_format(vec))
                round_poly = round_poly.add(rip)
            t.commit(round_poly.to_bytes())
            r = t.evaluate_challenge_into_field()
            cur = [p.partial_evaluation(r, 0) for p in cur]
            challenges.append(r)
            round_polys.append(round_poly)
        return ComposedSumcheckProof(round_polys, s % R_MOD), challenges


class MultiComposedSumcheckVerifier:
    @staticmethod
    def verify(polys, proof):  # :126-142
        t = FiatShamirTranscript()
        t.commit(composed_poly_to_bytes(polys))
        sub = MultiComposedSumcheckVerifier.verify_internal(proof, t)
        if sub is None:
            raise ValueError("Verification failed")
        tot = 0
        for p in polys:
            tot = (tot + p.evaluation(sub.challenges)) % R_MOD
        return tot == sub.sum

    @staticmethod
    def verify_partial(proof):  # :143-149
        t = FiatShamirTranscript()
        sub = MultiComposedSumcheckVerifier.verify_internal(proof, t)
        if sub is None:
            raise ValueError("Verification failed")
        return sub

    @staticmethod
    def verify_internal(proof, t):  # :151-181 ; returns None where the reference returns Err
        t.commit(be32(proof.sum))
        claimed = proof.sum
        challenges = []
        for rp in proof.round_polys:
            t.commit(rp.to_bytes())
            c = t.evaluate_challenge_into_field()
            challenges.append(c)
            if claimed != (rp.evaluate(0) + rp.evaluate(1)) % R_MOD:
                return None
            claimed = rp.evaluate(c)
        return SubClaim(claimed, challenges)


# --------------------------------------------------------------------------------------------
# Seeded synthetic inputs (the reference's generate_random_numbers uses an unseeded thread_rng,
# polynomial/src/utils.rs:256-259, so the convention below is this repo's own; it is restated in
# oracle/zkref.c and zk_cryptography_b200/csrc/ and all three must agree bit for bit).
# entry(seed, table, i): four splitmix64 words w0..w3 of the counter (seed, table, i, limb) form a
# 256-bit little-endian-limb integer; the canonical value is that integer mod r.
# --------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def synth_entry(seed, table, i):
    base = _splitmix64((seed & _M64) ^ _splitmix64((table * 0xD1342543DE82EF95 + 0x632BE59BD9B4E019) & _M64))
    v = 0
    for limb in range(4):
        w = _splitmix64((base + ((i * 4 + limb) * 0x9E3779B97F4A7C15)) & _M64)
        v |= w << (64 * limb)
    return v % R_MOD


def synth_table(seed, table, n_vars):
    return Multilinear([synth_entry(seed, table, i) for i in range(1 << n_vars)])
