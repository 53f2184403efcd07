"""TEST INFRASTRUCTURE ONLY -- Python big-int restatement of the reference's GKR layer driver and circuit model
(SURVEY.md 8f next-1), on top of oracle/pymodel.py.  Only tests/ and bench.py's CPU legs may import it.

Two provers that must agree byte for byte:
  * GKRProtocol.prove        -- the literal restatement: dense 2^(3i+2)-entry wiring tables, partial_evaluations over the
                                gate-label variables, add_distinct / mul_distinct, prove_partial (pure Python: small circuits)
  * GKRProtocol.prove_sparse -- the same values from the closed form add(r, b, c) = sum over add gates g of eq(r, g)
                                [b = in0(g), c = in1(g)], layer sumchecks through a caller-supplied prover (the C oracle
                                oracle/zkref.c for sizes where pure Python is too slow)
PARITY STATUS: "parity unpinned" for proof bytes (the reference's GKR tests only assert verify == true,
gkr/src/protocol.rs:209-285); pinned: the circuit-evaluation and wiring-table known answers of
circuit/src/circuit.rs:139-518 and circuit/src/utils.rs:38-64 (tests/test_oracle_kat.py).

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
from . import pymodel as pm

R = pm.R_MOD
ADD, MUL = "Add", "Mul"          # circuit/src/gate.rs:2-5


class Gate:                      # circuit/src/gate.rs:8-17
    def __init__(self, gate_type, inputs):
        self.gate_type, self.inputs = gate_type, list(inputs)


class CircuitLayer:              # circuit/src/circuit.rs:10-13, 21-25
    def __init__(self, layer):
        self.layer = list(layer)


def size_of_mle_n_var_at_each_layer(layer_index):   # circuit/src/utils.rs:1-10
    if layer_index == 0:
        return 1 << 3
    return 1 << (layer_index + 2 * (layer_index + 1))


def binary_string(index, bit_count):                # circuit/src/utils.rs:28-34
    if bit_count == 0:
        bit_count = 1
    b = format(index, "b")
    return "0" * max(bit_count - len(b), 0) + b


def transform_label_to_binary_and_to_decimal(layer_index, a, b, c):   # circuit/src/utils.rs:12-25
    s = binary_string(a, layer_index) + binary_string(b, layer_index + 1) + binary_string(c, layer_index + 1)
    return int(s, 2)


class Circuit:                   # circuit/src/circuit.rs:15-122
    def __init__(self, layers):
        self.layers = list(layers)

    def evaluation(self, inp):   # :32-55
        layers = [list(inp)]
        cur = list(inp)
        for layer in reversed(self.layers):
            cur = [(cur[g.inputs[0]] + cur[g.inputs[1]]) % R if g.gate_type == ADD else (cur[g.inputs[0]] * cur[g.inputs[1]]) % R for g in layer.layer]
            layers.append(cur)
        layers.reverse()
        return layers

    def add_mult_mle(self, layer_index):   # :57-95
        n = size_of_mle_n_var_at_each_layer(layer_index)
        add, mul = [0] * n, [0] * n
        for gi, g in enumerate(self.layers[layer_index].layer):
            d = transform_label_to_binary_and_to_decimal(layer_index, gi, g.inputs[0], g.inputs[1])
            (add if g.gate_type == ADD else mul)[d] = 1
        return pm.Multilinear(add), pm.Multilinear(mul)

    @staticmethod
    def random(num_of_layers):   # :97-121 (deterministic despite the name)
        layers = []
        for li in range(num_of_layers):
            n_in = 2 ** (li + 1)
            layers.append(CircuitLayer([Gate(ADD if li % 2 == 0 else MUL, [(g * 2) % n_in, (g * 2 + 1) % n_in]) for g in range(2 ** li)]))
        return Circuit(layers)


class GKRProof:                  # gkr/src/protocol.rs:10-15
    def __init__(self, sumcheck_proofs, wb_s, wc_s, w_0_mle):
        self.sumcheck_proofs, self.wb_s, self.wc_s, self.w_0_mle = sumcheck_proofs, wb_s, wc_s, w_0_mle

    def to_bytes(self):
        """Not in the reference (GKRProof has no serialiser): w_0, then per layer the sumcheck proof bytes, wb, wc.
        Used only to compare provers."""
        out = self.w_0_mle.to_bytes()
        for p, wb, wc in zip(self.sumcheck_proofs, self.wb_s, self.wc_s):
            out += p.to_bytes() + pm.be32(wb) + pm.be32(wc)
        return out


def eq_vector(r):
    """eq(r, a) for a in 0..2^len(r), a's most significant bit paired with r[0]: what len(r) folds of variable 0
    (Multilinear::partial_evaluations(r, [0; len]), evaluation_form.rs:143-159) leave of the indicator of a."""
    v = [1]
    for x in r:
        v = [e * f % R for e in v for f in ((1 - x) % R, x % R)]
    return v


def wiring_sparse(circuit, layer_index, r, scale=1):
    """The tables add(r, b, c), mul(r, b, c) of circuit layer `layer_index` as {index of (b, c): value}, times `scale`."""
    bits = layer_index + 1
    a_bits = max(layer_index, 1)
    eq = eq_vector(r)
    assert len(eq) == 1 << a_bits
    add, mul = {}, {}
    for gi, g in enumerate(circuit.layers[layer_index].layer):
        d = (g.inputs[0] << bits) | g.inputs[1]
        tgt = add if g.gate_type == ADD else mul
        tgt[d] = (tgt.get(d, 0) + eq[gi] * scale) % R
    return add, mul


def _dense(sparse, n):
    t = [0] * n
    for k, v in sparse.items():
        t[k] = v
    return t


def _python_layer_prover(tables, claimed_sum):
    """tables: [add_ab, wb+wc, mul_ab, wb*wc] as int lists -> (proof, challenges) by the pure-Python prover"""
    polys = [pm.ComposedMultilinear([pm.Multilinear(tables[0]), pm.Multilinear(tables[1])]),
             pm.ComposedMultilinear([pm.Multilinear(tables[2]), pm.Multilinear(tables[3])])]
    return pm.MultiComposedSumcheckProver.prove_partial(polys, claimed_sum)


class _RawProof:
    def __init__(self, b):
        self._b = b

    def to_bytes(self):
        return self._b


def c_layer_prover(tables, claimed_sum):
    """layer sumcheck through the C oracle (oracle/zkref.c), for sizes where the pure-Python prover is too slow"""
    import numpy as np

    from . import cref
    n = len(tables[0]).bit_length() - 1
    tabs = np.concatenate([cref.ints_to_canon(t) for t in tables])
    b, ch = cref.prove(2, n, [2, 2], tabs, claimed_sum)
    return _RawProof(b), ch


def c_evaluate(w, pts):
    from . import cref
    return cref.evaluation(w, pts)


class GKRProtocol:               # gkr/src/protocol.rs:17-195
    @staticmethod
    def prove(circuit, circuit_evaluation):   # :21-113 -- literal
        t = pm.FiatShamirTranscript()
        proofs, wb_s, wc_s = [], [], []
        w_0 = pm.Multilinear(list(circuit_evaluation[0]) + [0])                     # :31-34
        t.commit(w_0.to_bytes())                                                    # :35
        n_r = t.evaluate_n_challenge_into_field(w_0.n_vars)                         # :37
        claimed = w_0.evaluation(n_r)                                               # :38
        add1, mul1 = circuit.add_mult_mle(0)                                        # :40
        w1 = pm.Multilinear(circuit_evaluation[1])
        # generate_layer_one_prove_sumcheck, gkr/src/utils.rs:12-57
        add_rbc = add1.partial_evaluations(n_r, [0] * len(n_r))
        mul_rbc = mul1.partial_evaluations(n_r, [0] * len(n_r))
        polys = [pm.ComposedMultilinear([add_rbc, w1.add_distinct(w1)]), pm.ComposedMultilinear([mul_rbc, w1.mul_distinct(w1)])]
        proof, ch = pm.MultiComposedSumcheckProver.prove_partial(polys, claimed)
        t.commit(proof.to_bytes())
        proofs.append(proof)
        b, c = ch[:len(ch) // 2], ch[len(ch) // 2:]
        wb, wc = w1.evaluation(b), w1.evaluation(c)
        wb_s.append(wb); wc_s.append(wc)
        alpha, beta = t.evaluate_challenge_into_field(), t.evaluate_challenge_into_field()
        claimed = (alpha * wb + beta * wc) % R
        r_b, r_c = b, c
        for li in range(2, len(circuit_evaluation)):                                 # :65-105
            add, mul = circuit.add_mult_mle(li - 1)
            z = [0] * len(r_b)
            add_ab = add.partial_evaluations(r_b, z).scalar_mul(alpha).add(add.partial_evaluations(r_c, z).scalar_mul(beta))
            mul_ab = mul.partial_evaluations(r_b, z).scalar_mul(alpha).add(mul.partial_evaluations(r_c, z).scalar_mul(beta))
            w = pm.Multilinear(circuit_evaluation[li])
            polys = [pm.ComposedMultilinear([add_ab, w.add_distinct(w)]), pm.ComposedMultilinear([mul_ab, w.mul_distinct(w)])]
            proof, ch = pm.MultiComposedSumcheckProver.prove_partial(polys, claimed)
            t.commit(proof.to_bytes())
            proofs.append(proof)
            b, c = ch[:len(ch) // 2], ch[len(ch) // 2:]
            wb, wc = w.evaluation(b), w.evaluation(c)
            wb_s.append(wb); wc_s.append(wc)
            r_b, r_c = b, c
            alpha, beta = t.evaluate_challenge_into_field(), t.evaluate_challenge_into_field()
            claimed = (alpha * wb + beta * wc) % R
        out = GKRProof(proofs, wb_s, wc_s, w_0)
        out.last_b, out.last_c = r_b, r_c
        return out

    @staticmethod
    def prove_sparse(circuit, circuit_evaluation, layer_prover=_python_layer_prover, evaluate=None):
        """Same transcript, same bytes; wiring tables from the closed form, layer sumchecks through `layer_prover`
        (tables as int lists, claimed sum) -> (proof with .to_bytes(), challenges)."""
        evaluate = evaluate or (lambda w, pts: pm.Multilinear(w).evaluation(pts))
        t = pm.FiatShamirTranscript()
        proofs, wb_s, wc_s = [], [], []
        w_0 = pm.Multilinear(list(circuit_evaluation[0]) + [0])
        t.commit(w_0.to_bytes())
        n_r = t.evaluate_n_challenge_into_field(w_0.n_vars)
        claimed = w_0.evaluation(n_r)
        alpha, beta, r_b, r_c = 1, 0, n_r, None
        for li in range(1, len(circuit_evaluation)):
            w = [int(v) % R for v in circuit_evaluation[li]]
            n = len(w) * len(w)
            add, mul = wiring_sparse(circuit, li - 1, r_b, alpha)
            if r_c is not None:
                add_c, mul_c = wiring_sparse(circuit, li - 1, r_c, beta)
                for k, v in add_c.items():
                    add[k] = (add.get(k, 0) + v) % R
                for k, v in mul_c.items():
                    mul[k] = (mul.get(k, 0) + v) % R
            tables = [_dense(add, n), [(x + y) % R for x in w for y in w], _dense(mul, n), [x * y % R for x in w for y in w]]
            proof, ch = layer_prover(tables, claimed)
            t.commit(proof.to_bytes())
            proofs.append(proof)
            b, c = ch[:len(ch) // 2], ch[len(ch) // 2:]
            wb, wc = evaluate(w, b), evaluate(w, c)
            wb_s.append(wb); wc_s.append(wc)
            r_b, r_c = b, c
            alpha, beta = t.evaluate_challenge_into_field(), t.evaluate_challenge_into_field()
            claimed = (alpha * wb + beta * wc) % R
        out = GKRProof(proofs, wb_s, wc_s, w_0)
        out.last_b, out.last_c = r_b, r_c
        return out

    @staticmethod
    def verify(circuit, inp, proof):          # :115-195
        if len(proof.sumcheck_proofs) != len(proof.wb_s) or len(proof.sumcheck_proofs) != len(proof.wc_s):
            return False
        t = pm.FiatShamirTranscript()
        t.commit(proof.w_0_mle.to_bytes())
        n_r = t.evaluate_n_challenge_into_field(proof.w_0_mle.n_vars)
        claimed = proof.w_0_mle.evaluation(n_r)
        r_b, r_c, alpha, beta = [], [], 0, 0
        add1, mul1 = circuit.add_mult_mle(0)
        # generate_layer_one_verify_sumcheck, gkr/src/utils.rs:59-98
        p0 = proof.sumcheck_proofs[0]
        if claimed != p0.sum:
            return False
        t.commit(p0.to_bytes())
        sub = pm.MultiComposedSumcheckVerifier.verify_partial(p0)
        if sub is None:
            return False
        rbc = list(n_r) + list(sub.challenges)
        wb, wc = proof.wb_s[0], proof.wc_s[0]
        if (add1.evaluation(rbc) * ((wb + wc) % R) + mul1.evaluation(rbc) * (wb * wc % R)) % R != sub.sum:
            return False
        alpha, beta = t.evaluate_challenge_into_field(), t.evaluate_challenge_into_field()
        claimed = (alpha * wb + beta * wc) % R
        ch = sub.challenges
        r_b, r_c = ch[:len(ch) // 2], ch[len(ch) // 2:]       # (the reference leaves r_b, r_c empty for a 1-layer proof: :135-136)
        if len(proof.sumcheck_proofs) == 1:
            r_b, r_c = [], []
        for i in range(1, len(proof.sumcheck_proofs)):         # :155-181
            p = proof.sumcheck_proofs[i]
            if claimed != p.sum:
                return False
            t.commit(p.to_bytes())
            sub = pm.MultiComposedSumcheckVerifier.verify_partial(p)
            if sub is None:
                return False
            ch = sub.challenges
            r_b, r_c = ch[:len(ch) // 2], ch[len(ch) // 2:]
            wb, wc = proof.wb_s[i], proof.wc_s[i]
            alpha, beta = t.evaluate_challenge_into_field(), t.evaluate_challenge_into_field()
            claimed = (alpha * wb + beta * wc) % R
        w_in = pm.Multilinear(list(inp))
        return claimed == (alpha * w_in.evaluation(r_b) + beta * w_in.evaluation(r_c)) % R     # :183-192


# ---- layered circuits of any power-of-two widths (the product's zksc_gkr_prove_linear; BASELINE config 4 as written) ------------
# The reference's Circuit can only hold the pyramid (layer i: 2^i gates).  The protocol itself does not care: the label widths
# follow from the layers.  `prove_layered` is GKRProtocol::prove with DENSE width^2 layer tables (the reference's form) on a
# circuit given as per-layer gate lists; on a pyramid it is prove_sparse, byte for byte (tests/test_oracle_gkr_layered.py).
class LayeredCircuit:
    def __init__(self, log_width, layers):
        """log_width[i] = log2(gates of layer i), [-1] = log2(inputs); layers[i] = list of (type 0/1, in0, in1)"""
        self.log_width, self.layers = list(log_width), [list(l) for l in layers]
        for i, l in enumerate(self.layers):
            assert len(l) == 1 << self.log_width[i]
            assert all(t in (0, 1) and 0 <= a < (1 << self.log_width[i + 1]) and 0 <= b < (1 << self.log_width[i + 1]) for t, a, b in l)

    @staticmethod
    def from_circuit(circuit):
        return LayeredCircuit(list(range(len(circuit.layers) + 1)),
                              [[(0 if g.gate_type == ADD else 1, g.inputs[0], g.inputs[1]) for g in layer.layer] for layer in circuit.layers])

    def evaluation(self, inp):      # circuit/src/circuit.rs:32-55
        layers = [[int(v) % R for v in inp]]
        for l in reversed(self.layers):
            cur = layers[0]
            layers.insert(0, [(cur[a] * cur[b]) % R if t else (cur[a] + cur[b]) % R for t, a, b in l])
        return layers


def _layered_wiring(lc, li, points_and_scales):
    """{(b, c) index: value} of alpha add(r_b, b, c) + beta add(r_c, b, c), and of mul"""
    bits = lc.log_width[li + 1]
    wgt = None
    for r, scale in points_and_scales:
        e = [x * scale % R for x in eq_vector(r)]
        wgt = e if wgt is None else [(x + y) % R for x, y in zip(wgt, e)]
    assert len(wgt) >= len(lc.layers[li])
    add, mul = {}, {}
    for gi, (t, a, b) in enumerate(lc.layers[li]):
        d = (a << bits) | b
        tgt = mul if t else add
        tgt[d] = (tgt.get(d, 0) + wgt[gi]) % R
    return add, mul


def prove_layered(lc, evaluation, layer_prover=_python_layer_prover, evaluate=None):
    """GKRProtocol::prove (gkr/src/protocol.rs:21-113) on a LayeredCircuit, dense layer tables of width^2 entries"""
    evaluate = evaluate or (lambda w, pts: pm.Multilinear(w).evaluation(pts))
    t = pm.FiatShamirTranscript()
    proofs, wb_s, wc_s = [], [], []
    w0 = [int(v) % R for v in evaluation[0]]
    w_0 = pm.Multilinear(w0 + [0] if len(w0) == 1 else w0)                          # protocol.rs:31-34
    t.commit(w_0.to_bytes())
    r_b = t.evaluate_n_challenge_into_field(w_0.n_vars)
    claimed = w_0.evaluation(r_b)
    points = [(r_b, 1)]
    for li in range(len(lc.layers)):
        w = [int(v) % R for v in evaluation[li + 1]]
        n = len(w) * len(w)
        add, mul = _layered_wiring(lc, li, points)
        tables = [_dense(add, n), [(x + y) % R for x in w for y in w], _dense(mul, n), [x * y % R for x in w for y in w]]
        proof, ch = layer_prover(tables, claimed)
        t.commit(proof.to_bytes())
        proofs.append(proof)
        b, c = ch[:len(ch) // 2], ch[len(ch) // 2:]
        wb, wc = evaluate(w, b), evaluate(w, c)
        wb_s.append(wb); wc_s.append(wc)
        alpha, beta = t.evaluate_challenge_into_field(), t.evaluate_challenge_into_field()
        claimed = (alpha * wb + beta * wc) % R
        points = [(b, alpha), (c, beta)]
    out = GKRProof(proofs, wb_s, wc_s, w_0)
    out.last_b, out.last_c = b, c
    return out


class SuccintGKRProof:           # gkr/src/succint_protocol.rs:21-29
    def __init__(self, gkr, proof_wb_opening, proof_wc_opening):
        self.sumcheck_proofs, self.wb_s, self.wc_s, self.w_0_mle = gkr.sumcheck_proofs, gkr.wb_s, gkr.wc_s, gkr.w_0_mle
        self.proof_wb_opening, self.proof_wc_opening = proof_wb_opening, proof_wc_opening      # (evaluation, [G1 proofs]) or None


class SuccintGKRProtocol:        # gkr/src/succint_protocol.rs:35-266
    @staticmethod
    def prove(circuit, circuit_evaluation, tau):   # :37-167; tau: oracle.kzgmodel.TrustedSetup
        """The transcript and the layer sumchecks are GKRProtocol::prove's (the two functions differ only after the last layer's
        sumcheck); then the input layer, blown up to the trusted setup's size by add_to_back, is committed and opened at
        (b, 0, ...) and (c, 0, ...)  (:136-157).  A circuit with a single layer never enters the loop: default commitment / proofs."""
        from oracle import kzgmodel as k
        base = GKRProtocol.prove_sparse(circuit, circuit_evaluation)
        if len(circuit_evaluation) < 3:
            return None, SuccintGKRProof(base, None, None)
        w = [int(v) % R for v in circuit_evaluation[-1]]
        exponent = len(tau.powers_of_tau_in_g1).bit_length() - 1                       # gkr/src/utils.rs:100-111
        blow = exponent - (len(w).bit_length() - 1)
        poly = [v for v in w for _ in range(1 << blow)]                                # add_to_back, evaluation_form.rs:98-110
        b = list(base.last_b) + [0] * (exponent - len(base.last_b))
        c = list(base.last_c) + [0] * (exponent - len(base.last_c))
        return k.commitment(poly, tau), SuccintGKRProof(base, k.open_(poly, b, tau), k.open_(poly, c, tau))

    @staticmethod
    def verify(circuit, commitment, proof, tau):   # :169-266, the two pairing checks restated on the group elements (tau is known here)
        from oracle import kzgmodel as k
        if len(proof.sumcheck_proofs) != len(proof.wb_s) or len(proof.sumcheck_proofs) != len(proof.wc_s):
            return False
        t = pm.FiatShamirTranscript()
        t.commit(proof.w_0_mle.to_bytes())
        n_r = t.evaluate_n_challenge_into_field(proof.w_0_mle.n_vars)
        claimed = proof.w_0_mle.evaluation(n_r)
        p0 = proof.sumcheck_proofs[0]
        if claimed != p0.sum:                                                          # generate_layer_one_verify_sumcheck, utils.rs:59-98
            return False
        t.commit(p0.to_bytes())
        sub = pm.MultiComposedSumcheckVerifier.verify_partial(p0)
        add1, mul1 = circuit.add_mult_mle(0)
        rbc = list(n_r) + list(sub.challenges)
        wb, wc = proof.wb_s[0], proof.wc_s[0]
        if (add1.evaluation(rbc) * ((wb + wc) % R) + mul1.evaluation(rbc) * (wb * wc % R)) % R != sub.sum:
            return False
        alpha, beta = t.evaluate_challenge_into_field(), t.evaluate_challenge_into_field()
        claimed = (alpha * wb + beta * wc) % R
        r_b, r_c = [], []
        for i in range(1, len(proof.sumcheck_proofs)):
            p = proof.sumcheck_proofs[i]
            if claimed != p.sum:
                return False
            t.commit(p.to_bytes())
            alpha, beta = t.evaluate_challenge_into_field(), t.evaluate_challenge_into_field()
            ch = pm.MultiComposedSumcheckVerifier.verify_partial(p).challenges
            r_b, r_c = ch[:len(ch) // 2], ch[len(ch) // 2:]
            claimed = (alpha * proof.wb_s[i] + beta * proof.wc_s[i]) % R
        n_tau = len(tau.tau)
        rb = list(r_b) + [0] * (n_tau - len(r_b))
        rc = list(r_c) + [0] * (n_tau - len(r_c))
        ok_b = proof.proof_wb_opening is not None and k.verify_group(commitment, rb, proof.proof_wb_opening, tau)
        ok_c = proof.proof_wc_opening is not None and k.verify_group(commitment, rc, proof.proof_wc_opening, tau)
        eb, ec = (proof.proof_wb_opening[0], proof.proof_wc_opening[0]) if (ok_b and ok_c) else (0, 0)
        return claimed == (alpha * eb + beta * ec) % R
