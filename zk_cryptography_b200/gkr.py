"""GKR layer driver over the zksc C ABI: host-side mirror of the reference's `circuit` and `gkr` crates.

  Gate, GateType, CircuitLayer, Circuit      circuit/src/{gate.rs, circuit.rs, utils.rs}
  GKRProof, GKRProtocol                      gkr/src/{protocol.rs, utils.rs}

Per layer the reference builds four 2^(2k)-entry tables (k = bits of a gate input label) and proves
    sum_{b,c} add~(b,c) (W(b) + W(c)) + mul~(b,c) (W(b) W(c))           (gkr/src/protocol.rs:67-93)
with MultiComposedSumcheckProver::prove_partial.  Here the tables are built ON THE DEVICE, straight into the
prover's table handle:
  * W(b) + W(c), W(b) W(c)   (add_distinct / mul_distinct, evaluation_form.rs:28-52): one outer-sum / outer-product
    kernel each from the 2^k layer values                                        -> zksc_tables_fill_outer
  * add~ = alpha add(r_b, ., .) + beta add(r_c, ., .) (and mul~): the reference folds a dense 0/1 table of 2^(3k-1)
    entries k-1 times; what is left is  sum over gates g of eq(r, g) at entry (in0(g), in1(g))  -- at most one
    entry per gate.  Those few values are formed on the host (exact field arithmetic, hence bit-identical) and
    scattered into a zeroed table                                                -> zksc_tables_fill_sparse
  * the layer sumcheck (P = 2, d = (2, 2)) runs in the CUDA round kernels        -> zksc_prove
  * W(b*), W(c*) by k folds on the device                                        -> zksc_ml_evaluation
The Fiat-Shamir transcripts (outer GKR transcript and the per-layer sumcheck transcripts) stay on the host.
There is no CPU fallback for the table-sized work.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import PROTO_MULTI_PARTIAL, Tables, ZkscError, from_mont, to_mont
from .api import FiatShamirTranscript, Multilinear, MultiComposedProof, MultiComposedSumcheckVerifier, SparseUnivariatePolynomial, default_context

R = _lib.R_MOD


class GateType:                  # circuit/src/gate.rs:2-5
    Add = "Add"
    Mul = "Mul"


class Gate:                      # circuit/src/gate.rs:8-17
    def __init__(self, gate_type, inputs):
        self.gate_type, self.inputs = gate_type, list(inputs)

    new = classmethod(lambda cls, gate_type, inputs: cls(gate_type, inputs))


class CircuitLayer:              # circuit/src/circuit.rs:10-25
    def __init__(self, layer):
        self.layer = list(layer)

    new = classmethod(lambda cls, layer: cls(layer))


def size_of_mle_n_var_at_each_layer(layer_index):   # circuit/src/utils.rs:1-10
    return 1 << 3 if layer_index == 0 else 1 << (layer_index + 2 * (layer_index + 1))


def transform_label_to_binary_and_to_decimal(layer_index, a, b, c):   # circuit/src/utils.rs:12-34
    bits = layer_index + 1
    for v, w in ((a, max(layer_index, 1)), (b, bits), (c, bits)):
        if v >> w:
            raise ZkscError(-3, "gate label does not fit its bit field")
    return (a << (2 * bits)) | (b << bits) | c


class Circuit:                   # circuit/src/circuit.rs:15-122
    def __init__(self, layers):
        self.layers = list(layers)

    new = classmethod(lambda cls, layers: cls(layers))

    def evaluation(self, inp):   # :32-55 -- one pass over the gates, host (not table-sized work of the sumcheck path)
        layers = [[int(v) % R for v in inp]]
        cur = layers[0]
        for layer in reversed(self.layers):
            cur = [(cur[g.inputs[0]] + cur[g.inputs[1]]) % R if g.gate_type == GateType.Add else (cur[g.inputs[0]] * cur[g.inputs[1]]) % R
                   for g in layer.layer]
            layers.append(cur)
        layers.reverse()
        return layers

    def add_mult_mle(self, layer_index):   # :57-95 (dense; the prover below never materialises it beyond layer 0)
        n = size_of_mle_n_var_at_each_layer(layer_index)
        add, mul = [0] * n, [0] * n
        for gi, g in enumerate(self.layers[layer_index].layer):
            d = transform_label_to_binary_and_to_decimal(layer_index, gi, g.inputs[0], g.inputs[1])
            (add if g.gate_type == GateType.Add else mul)[d] = 1
        return Multilinear(add), Multilinear(mul)

    @staticmethod
    def random(num_of_layers):   # :97-121
        layers = []
        for li in range(num_of_layers):
            n_in = 2 ** (li + 1)
            gt = GateType.Add if li % 2 == 0 else GateType.Mul
            layers.append(CircuitLayer([Gate(gt, [(g * 2) % n_in, (g * 2 + 1) % n_in]) for g in range(2 ** li)]))
        return Circuit(layers)


class GKRProof:                  # gkr/src/protocol.rs:10-15
    def __init__(self, sumcheck_proofs, wb_s, wc_s, w_0_mle):
        self.sumcheck_proofs, self.wb_s, self.wc_s, self.w_0_mle = sumcheck_proofs, wb_s, wc_s, w_0_mle

    def to_bytes(self):
        """w_0, then per layer the sumcheck proof bytes (ComposedSumcheckProof::to_bytes), wb, wc.  The reference has no
        GKRProof serialiser; this is what its transcript absorbs plus the two claimed evaluations, for comparisons."""
        out = b"".join(int(e).to_bytes(32, "big") for e in self.w_0_mle.to_ints())
        for p, wb, wc in zip(self.sumcheck_proofs, self.wb_s, self.wc_s):
            out += p.to_bytes() + int(wb).to_bytes(32, "big") + int(wc).to_bytes(32, "big")
        return out


def _eq_vector(r):
    """eq(r, a), a = 0..2^len(r) - 1 with a's most significant bit paired with r[0]: what partial_evaluations(r, [0; len])
    (evaluation_form.rs:143-159) leaves of the indicator of a."""
    v = [1]
    for x in r:
        nx = (1 - x) % R
        v = [e * f % R for e in v for f in (nx, x)]
    return v


def _wiring_entries(circuit, layer_index, points_and_scales):
    """{(in0 << bits) | in1: value} of  sum_j scale_j * add(r_j, ., .)  and the same for mul, circuit layer `layer_index`."""
    bits = layer_index + 1
    add, mul = {}, {}
    for r, scale in points_and_scales:
        eq = _eq_vector(r)
        if len(eq) != 1 << max(layer_index, 1):
            raise ZkscError(-3, "challenge vector does not match the layer's gate-label bits")
        for gi, g in enumerate(circuit.layers[layer_index].layer):
            if (g.inputs[0] >> bits) or (g.inputs[1] >> bits):
                raise ZkscError(-3, "gate input label does not fit its bit field")
            d = (g.inputs[0] << bits) | g.inputs[1]
            tgt = add if g.gate_type == GateType.Add else mul
            tgt[d] = (tgt.get(d, 0) + eq[gi] * scale) % R
    return add, mul


def _fill_sparse(tables, index, entries):
    idx = np.fromiter(entries.keys(), dtype=np.uint64, count=len(entries))
    vals = to_mont(list(entries.values())) if entries else np.zeros((0, 4), dtype=np.uint64)
    tables.fill_sparse(index, idx, vals)


class GKRInstance:
    """The arguments of zksc_gkr_prove in the form a Rust caller already holds them (`&Circuit`, `&Vec<Vec<F>>` with F in
    Montgomery limbs): prepared once, so that `prove_raw` is nothing but the C call."""

    def __init__(self, circuit, circuit_evaluation):
        L = self.n_layers = len(circuit.layers)
        if len(circuit_evaluation) != L + 1:
            raise ZkscError(-3, "circuit_evaluation must hold one vector per layer plus the input")
        gates = [g for l in circuit.layers for g in l.layer]
        for g in gates:
            if len(g.inputs) != 2 or min(g.inputs) < 0 or max(g.inputs) >= 1 << 32:
                raise ZkscError(-3, "a gate takes two input labels")
        self.n_gates = np.asarray([len(l.layer) for l in circuit.layers], dtype=np.uint32)
        self.gtype = np.asarray([0 if g.gate_type == GateType.Add else 1 for g in gates], dtype=np.uint8)
        self.in0 = np.asarray([g.inputs[0] for g in gates], dtype=np.uint32)
        self.in1 = np.asarray([g.inputs[1] for g in gates], dtype=np.uint32)
        self.vals = [to_mont([int(v) % R for v in layer]) for layer in circuit_evaluation]
        self.vlen = np.asarray([len(layer) for layer in circuit_evaluation], dtype=np.uint64)
        self.ptrs = (_lib._u64p * len(self.vals))(*[_lib.p64(v) for v in self.vals])
        self.rounds = int(_lib.lib().zksc_gkr_total_rounds(L))

    def prove_raw(self, ctx):
        L, rounds = self.n_layers, self.rounds
        out = dict(w0=np.zeros((2, 4), dtype=np.uint64), sums=np.zeros((L, 4), dtype=np.uint64), wb=np.zeros((L, 4), dtype=np.uint64),
                   wc=np.zeros((L, 4), dtype=np.uint64), msgs=np.zeros((rounds, 6, 4), dtype=np.uint64), lens=np.zeros(rounds, dtype=np.uint32),
                   chal=np.zeros((rounds, 4), dtype=np.uint64))
        ctx.check(_lib.lib().zksc_gkr_prove(ctx._h, L, _lib.p32(self.n_gates), _lib.p8(self.gtype), _lib.p32(self.in0), _lib.p32(self.in1), self.ptrs,
                                            _lib.p64(self.vlen), _lib.p64(out["w0"]), _lib.p64(out["sums"]), _lib.p64(out["wb"]), _lib.p64(out["wc"]),
                                            _lib.p64(out["msgs"]), _lib.p32(out["lens"]), _lib.p64(out["chal"])))
        return out

    def parse(self, raw):
        msgs, lens = raw["msgs"], raw["lens"]
        proofs, off = [], 0
        sums_i, all_ints = from_mont(raw["sums"]), from_mont(msgs.reshape(-1, 4))
        for li in range(self.n_layers):
            n = 2 * (li + 1)
            rps = []
            for r in range(off, off + n):
                m = int(lens[r])
                v = all_ints[r * 6:r * 6 + 2 * m]
                rps.append(SparseUnivariatePolynomial([(v[2 * i], v[2 * i + 1]) for i in range(m)]))
            proofs.append(MultiComposedProof(rps, sums_i[li]))
            off += n
        proof = GKRProof(proofs, from_mont(raw["wb"]), from_mont(raw["wc"]), Multilinear(raw["w0"]))
        proof.challenges = from_mont(raw["chal"])
        return proof


class GKRProtocol:               # gkr/src/protocol.rs:17-195
    @staticmethod
    def _layer(ctx, circuit, layer_index, w_ints, points_and_scales, claimed_sum, transcript, proofs, wb_s, wc_s):
        """One layer: build the four tables on the device, prove_partial, absorb the proof, evaluate W at the two halves of
        the challenge vector, draw (alpha, beta).  gkr/src/utils.rs:12-57 and gkr/src/protocol.rs:66-104."""
        k = len(w_ints).bit_length() - 1
        if len(w_ints) != 1 << k:
            raise ZkscError(-3, "Number of evaluations must be a power of 2")
        w = to_mont(w_ints)
        add, mul = _wiring_entries(circuit, layer_index, points_and_scales)
        t = Tables.alloc(ctx, 2 * k, [2, 2])
        try:
            _fill_sparse(t, 0, add)                 # alpha add(r_b, b, c) + beta add(r_c, b, c)
            t.fill_outer(1, False, w, w)            # W(b) + W(c)      add_distinct
            _fill_sparse(t, 2, mul)                 # alpha mul(r_b, b, c) + beta mul(r_c, b, c)
            t.fill_outer(3, True, w, w)             # W(b) W(c)        mul_distinct
            msgs, lens, chal = t.prove(PROTO_MULTI_PARTIAL, to_mont(claimed_sum))
        finally:
            t.free()
        n = 2 * k
        rps = []
        for r in range(n):
            m = int(lens[0, r])
            v = from_mont(msgs[0, r, :2 * m]) if m else []
            rps.append(SparseUnivariatePolynomial([(v[2 * i], v[2 * i + 1]) for i in range(m)]))
        proof = MultiComposedProof(rps, claimed_sum % R)
        transcript.commit(proof.to_bytes())
        proofs.append(proof)
        ch = from_mont(chal[0]) if n else []
        b, c = ch[:len(ch) // 2], ch[len(ch) // 2:]
        wm = Multilinear(w)
        eval_wb, eval_wc = wm.evaluation(b), wm.evaluation(c)
        wb_s.append(eval_wb)
        wc_s.append(eval_wc)
        alpha, beta = transcript.evaluate_challenge_into_field(), transcript.evaluate_challenge_into_field()
        return (alpha * eval_wb + beta * eval_wc) % R, alpha, beta, b, c

    @staticmethod
    def prove(circuit, circuit_evaluation, ctx=None):   # :21-113
        """The whole proof in ONE C call (zksc_gkr_prove): layer loop, table construction, layer sumchecks, W(b*), W(c*) and
        the outer transcript all inside libzksc.so.  `prove_layerwise` below is the same protocol driven layer by layer
        from Python over the finer-grained C entry points; both must give the same bytes (tests/test_gpu_gkr.py)."""
        inst = GKRInstance(circuit, circuit_evaluation)
        return inst.parse(inst.prove_raw(ctx or default_context()))

    @staticmethod
    def prove_layerwise(circuit, circuit_evaluation, ctx=None):   # :21-113, one C call per table operation
        ctx = ctx or default_context()
        transcript = FiatShamirTranscript()
        proofs, wb_s, wc_s = [], [], []
        w_0_mle = Multilinear([int(v) % R for v in circuit_evaluation[0]] + [0])            # :31-34
        transcript.commit(b"".join(int(e).to_bytes(32, "big") for e in w_0_mle.to_ints()))  # :35
        n_r = transcript.evaluate_n_challenge_into_field(w_0_mle.n_vars)                    # :37
        claimed = w_0_mle.evaluation(n_r)                                                   # :38
        # layer one: add(n_r, b, c), mul(n_r, b, c) unscaled (gkr/src/utils.rs:23-33)
        claimed, alpha, beta, r_b, r_c = GKRProtocol._layer(ctx, circuit, 0, circuit_evaluation[1], [(n_r, 1)], claimed, transcript, proofs, wb_s, wc_s)
        for layer_index in range(2, len(circuit_evaluation)):                               # :65-105
            claimed, alpha, beta, r_b, r_c = GKRProtocol._layer(ctx, circuit, layer_index - 1, circuit_evaluation[layer_index],
                                                                [(r_b, alpha), (r_c, beta)], claimed, transcript, proofs, wb_s, wc_s)
        return GKRProof(proofs, wb_s, wc_s, w_0_mle)

    @staticmethod
    def verify(circuit, inp, proof):          # :115-195
        if len(proof.sumcheck_proofs) != len(proof.wb_s) or len(proof.sumcheck_proofs) != len(proof.wc_s):
            return False
        transcript = FiatShamirTranscript()
        transcript.commit(b"".join(int(e).to_bytes(32, "big") for e in proof.w_0_mle.to_ints()))
        n_r = transcript.evaluate_n_challenge_into_field(proof.w_0_mle.n_vars)
        claimed = proof.w_0_mle.evaluation(n_r)
        # generate_layer_one_verify_sumcheck, gkr/src/utils.rs:59-98
        p0 = proof.sumcheck_proofs[0]
        if claimed != p0.sum:
            return False
        transcript.commit(p0.to_bytes())
        try:
            sub = MultiComposedSumcheckVerifier.verify_partial(p0)
        except ZkscError as e:
            if e.code == -7:
                return False
            raise
        add1, mul1 = circuit.add_mult_mle(0)
        rbc = list(n_r) + list(sub.challenges)
        wb, wc = proof.wb_s[0], proof.wc_s[0]
        if (add1.evaluation(rbc) * ((wb + wc) % R) + mul1.evaluation(rbc) * (wb * wc % R)) % R != sub.sum:
            return False
        alpha, beta = transcript.evaluate_challenge_into_field(), transcript.evaluate_challenge_into_field()
        claimed = (alpha * wb + beta * wc) % R
        r_b, r_c = [], []
        for i in range(1, len(proof.sumcheck_proofs)):         # :155-181
            p = proof.sumcheck_proofs[i]
            if claimed != p.sum:
                return False
            transcript.commit(p.to_bytes())
            try:
                sub = MultiComposedSumcheckVerifier.verify_partial(p)
            except ZkscError as e:
                if e.code == -7:
                    return False
                raise
            ch = sub.challenges
            r_b, r_c = ch[:len(ch) // 2], ch[len(ch) // 2:]
            wb, wc = proof.wb_s[i], proof.wc_s[i]
            alpha, beta = transcript.evaluate_challenge_into_field(), transcript.evaluate_challenge_into_field()
            claimed = (alpha * wb + beta * wc) % R
        w_in = Multilinear([int(v) % R for v in inp])
        return claimed == (alpha * w_in.evaluation(r_b) + beta * w_in.evaluation(r_c)) % R     # :183-192


class LayeredCircuit:
    """A layered add/mul circuit whose layers have ANY power-of-two widths, resident on the device, proved in time linear in its gates
    (zksc_circuit_* / zksc_gkr_prove_linear, csrc/gkr_linear.cuh): BASELINE config 4 read literally ("width 2^20, depth 8").  The
    reference's `Circuit` holds only the pyramid shape (layer i: 2^i gates); on such circuits this prover gives the bytes of
    GKRProtocol::prove (gkr/src/protocol.rs:21-113), and its proofs are checked by `verify` below, GKRProtocol::verify (:115-195) with
    the label widths taken from the layers.

    log_width[i] = log2(gates of layer i), output layer first, log_width[-1] = log2(inputs); gate arrays of all layers concatenated:
    gtype (0 = Add, 1 = Mul), in0, in1 (wire indices of the layer below)."""

    def __init__(self, log_width, gtype, in0, in1, ctx=None):
        self.ctx = ctx or default_context()
        self.log_width = [int(x) for x in log_width]
        self.n_layers = len(self.log_width) - 1
        self.gtype = np.ascontiguousarray(gtype, dtype=np.uint8)
        self.in0 = np.ascontiguousarray(in0, dtype=np.uint32)
        self.in1 = np.ascontiguousarray(in1, dtype=np.uint32)
        n_gates = sum(1 << w for w in self.log_width[:-1])
        if not (len(self.gtype) == len(self.in0) == len(self.in1) == n_gates):
            raise ZkscError(-3, "gate arrays must hold 2^log_width[i] gates per layer")
        lw = np.asarray(self.log_width, dtype=np.uint32)
        h = ctypes.c_void_p()
        self.ctx.check(_lib.lib().zksc_circuit_create(self.ctx._h, self.n_layers, _lib.p32(lw), _lib.p8(self.gtype), _lib.p32(self.in0), _lib.p32(self.in1),
                                                      ctypes.byref(h)))
        self._h = h
        self.rounds = int(_lib.lib().zksc_circuit_total_rounds(self._h))

    @classmethod
    def from_circuit(cls, circuit, ctx=None):
        """the reference's pyramid `Circuit` (layer i: 2^i gates) in this form"""
        L = len(circuit.layers)
        gt, a, b = [], [], []
        for i, layer in enumerate(circuit.layers):
            if len(layer.layer) != 1 << i:
                raise ZkscError(-3, "layer %d must have 2^%d gates" % (i, i))
            for g in layer.layer:
                gt.append(0 if g.gate_type == GateType.Add else 1); a.append(g.inputs[0]); b.append(g.inputs[1])
        return cls(list(range(L + 1)), gt, a, b, ctx)

    @classmethod
    def random(cls, log_width, seed=0, ctx=None):
        """every gate of layer i: a uniformly random type and two uniformly random wires of layer i + 1"""
        rng = np.random.default_rng(seed)
        gt, a, b = [], [], []
        for i in range(len(log_width) - 1):
            n, m = 1 << log_width[i], 1 << log_width[i + 1]
            gt.append(rng.integers(0, 2, n, dtype=np.uint8)); a.append(rng.integers(0, m, n, dtype=np.uint32)); b.append(rng.integers(0, m, n, dtype=np.uint32))
        return cls(log_width, np.concatenate(gt), np.concatenate(a), np.concatenate(b), ctx)

    def layers(self):
        """-> per layer (gtype, in0, in1) array views"""
        out, off = [], 0
        for w in self.log_width[:-1]:
            n = 1 << w
            out.append((self.gtype[off:off + n], self.in0[off:off + n], self.in1[off:off + n]))
            off += n
        return out

    def evaluate(self, inputs, mont=False):
        """Circuit::evaluation (circuit/src/circuit.rs:32-55) on the device; -> the outputs (ints; with mont=True inputs and outputs are
        Montgomery limb arrays, as a Rust caller holds its Vec<Fr>); the layer values stay in HBM"""
        x = np.ascontiguousarray(inputs, dtype=np.uint64) if mont else to_mont([int(v) % R for v in inputs])
        if x.shape != (1 << self.log_width[-1], 4):
            raise ZkscError(-3, "the input layer has 2^%d values" % self.log_width[-1])
        out = np.zeros((1 << self.log_width[0], 4), dtype=np.uint64)
        self.ctx.check(_lib.lib().zksc_circuit_evaluate(self._h, _lib.p64(x), _lib.p64(out)))
        return out if mont else from_mont(out)

    def layer_values(self, layer):
        out = np.zeros((1 << self.log_width[layer], 4), dtype=np.uint64)
        self.ctx.check(_lib.lib().zksc_circuit_layer_values(self._h, layer, _lib.p64(out)))
        return from_mont(out)

    def prove_raw(self):
        L, rounds = self.n_layers, self.rounds
        out = dict(w0=np.zeros((max(2, 1 << self.log_width[0]), 4), dtype=np.uint64), sums=np.zeros((L, 4), dtype=np.uint64),
                   wb=np.zeros((L, 4), dtype=np.uint64), wc=np.zeros((L, 4), dtype=np.uint64), msgs=np.zeros((rounds, 6, 4), dtype=np.uint64),
                   lens=np.zeros(rounds, dtype=np.uint32), chal=np.zeros((rounds, 4), dtype=np.uint64))
        self.ctx.check(_lib.lib().zksc_gkr_prove_linear(self._h, _lib.p64(out["w0"]), _lib.p64(out["sums"]), _lib.p64(out["wb"]), _lib.p64(out["wc"]),
                                                        _lib.p64(out["msgs"]), _lib.p32(out["lens"]), _lib.p64(out["chal"])))
        return out

    def prove(self):
        """GKRProtocol::prove of the latest `evaluate` -> GKRProof (same classes as the dense prover's)"""
        raw = self.prove_raw()
        msgs, lens = raw["msgs"], raw["lens"]
        proofs, off = [], 0
        sums_i, all_ints = from_mont(raw["sums"]), from_mont(msgs.reshape(-1, 4))
        for li in range(self.n_layers):
            n = 2 * self.log_width[li + 1]
            rps = []
            for r in range(off, off + n):
                m = int(lens[r])
                v = all_ints[r * 6:r * 6 + 2 * m]
                rps.append(SparseUnivariatePolynomial([(v[2 * i], v[2 * i + 1]) for i in range(m)]))
            proofs.append(MultiComposedProof(rps, sums_i[li]))
            off += n
        proof = GKRProof(proofs, from_mont(raw["wb"]), from_mont(raw["wc"]), Multilinear(raw["w0"]))
        proof.challenges = from_mont(raw["chal"])
        return proof

    def wiring_at_device(self, layer, points_and_scales, b, c):
        """the same two values from the gate lists on the device (zksc_circuit_wiring_eval): what makes `verify` cheap at width 2^20"""
        (r_b, alpha) = points_and_scales[0]
        two = len(points_and_scales) > 1
        out = np.zeros((2, 4), dtype=np.uint64)
        null = ctypes.POINTER(ctypes.c_uint64)()
        rc_arr = to_mont([int(x) for x in points_and_scales[1][0]]) if two else None
        beta_arr = to_mont([int(points_and_scales[1][1])]) if two else None
        self.ctx.check(_lib.lib().zksc_circuit_wiring_eval(self._h, layer, _lib.p64(to_mont([int(x) for x in r_b])), _lib.p64(to_mont([int(alpha)])),
                                                           _lib.p64(rc_arr) if two else null, _lib.p64(beta_arr) if two else null,
                                                           _lib.p64(to_mont([int(x) for x in b])), _lib.p64(to_mont([int(x) for x in c])), _lib.p64(out)))
        v = from_mont(out)
        return v[0], v[1]

    def wiring_at(self, layer, points_and_scales, b, c):
        """(add~, mul~)(b, c) = sum over the layer's gates g of [sum_j scale_j eq(r_j, g)] eq(b, in0 g) eq(c, in1 g), on the host
        (what Multilinear::evaluation of the wiring tables gives, gkr/src/protocol.rs:131-133, 164-171)"""
        gt, i0, i1 = self.layers()[layer]
        wgt = None
        for r, scale in points_and_scales:
            e = [x * scale % R for x in _eq_vector(r)]
            wgt = e if wgt is None else [(x + y) % R for x, y in zip(wgt, e)]
        eb, ec = _eq_vector(b), _eq_vector(c)
        add = mul = 0
        for g in range(len(gt)):
            t = wgt[g] * eb[int(i0[g])] % R * ec[int(i1[g])] % R
            if gt[g]:
                mul = (mul + t) % R
            else:
                add = (add + t) % R
        return add, mul

    def verify(self, inputs, proof, device=True):
        """GKRProtocol::verify (gkr/src/protocol.rs:115-195) for this circuit: transcript, round checks and claims on the host; the wiring
        polynomials add~, mul~ at (r, b*, c*) and the input layer at (b*, c*) on the device (device=False: everything in Python integers,
        the independent cross-check of the tests, small circuits only)"""
        if len(proof.sumcheck_proofs) != self.n_layers or len(proof.wb_s) != self.n_layers or len(proof.wc_s) != self.n_layers:
            return False
        transcript = FiatShamirTranscript()
        transcript.commit(b"".join(int(e).to_bytes(32, "big") for e in proof.w_0_mle.to_ints()))
        r_b = transcript.evaluate_n_challenge_into_field(proof.w_0_mle.n_vars)
        claimed = proof.w_0_mle.evaluation(r_b)
        points = [(r_b, 1)]
        alpha = beta = r_c = None
        for i, p in enumerate(proof.sumcheck_proofs):
            if claimed != p.sum:
                return False
            transcript.commit(p.to_bytes())
            try:
                sub = MultiComposedSumcheckVerifier.verify_partial(p)
            except ZkscError as e:
                if e.code == -7:
                    return False
                raise
            ch = sub.challenges
            if len(ch) != 2 * self.log_width[i + 1]:
                return False
            b, c = ch[:len(ch) // 2], ch[len(ch) // 2:]
            wb, wc = proof.wb_s[i], proof.wc_s[i]
            add, mul = (self.wiring_at_device if device else self.wiring_at)(i, points, b, c)
            if (add * ((wb + wc) % R) + mul * (wb * wc % R)) % R != sub.sum:
                return False
            alpha, beta = transcript.evaluate_challenge_into_field(), transcript.evaluate_challenge_into_field()
            claimed = (alpha * wb + beta * wc) % R
            r_b, r_c = b, c
            points = [(r_b, alpha), (r_c, beta)]
        w_in = Multilinear(inputs) if isinstance(inputs, np.ndarray) else Multilinear([int(v) % R for v in inputs])   # (an array: Montgomery limbs)
        return claimed == (alpha * w_in.evaluation(r_b) + beta * w_in.evaluation(r_c)) % R

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            _lib.lib().zksc_circuit_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _eq_vector(r):
    """eq(r, a), a's most significant bit paired with r[0] (successive variable-0 folds, evaluation_form.rs:143-159)"""
    v = [1]
    for x in r:
        v = [e * f % R for e in v for f in ((1 - x) % R, x % R)]
    return v


class SuccintGKRProof:           # gkr/src/succint_protocol.rs:21-29
    def __init__(self, gkr, proof_wb_opening, proof_wc_opening):
        self.sumcheck_proofs, self.wb_s, self.wc_s, self.w_0_mle = gkr.sumcheck_proofs, gkr.wb_s, gkr.wc_s, gkr.w_0_mle
        self.proof_wb_opening, self.proof_wc_opening = proof_wb_opening, proof_wc_opening


class SuccintGKRProtocol:        # gkr/src/succint_protocol.rs:35-167
    @staticmethod
    def prove(circuit, circuit_evaluation, tau, ctx=None):
        """SuccintGKRProtocol::prove: GKRProtocol::prove's transcript and layer sumchecks (one C call, zksc_gkr_prove), then the input
        layer -- blown up to the trusted setup's size (add_to_back, evaluation_form.rs:98-110) -- committed and opened at (b, 0, ..)
        and (c, 0, ..) on the device (zksc_g1_msm, zksc_kzg_open; :136-157).  -> (commitment (18,) uint64 or None, SuccintGKRProof)."""
        from .kzg import MultilinearKZG
        ctx = ctx or default_context()
        base = GKRProtocol.prove(circuit, circuit_evaluation, ctx)
        if len(circuit_evaluation) < 3:       # the loop of :81 never runs: default commitment and openings
            return None, SuccintGKRProof(base, None, None)
        L = len(circuit.layers)
        ch = base.challenges[-2 * L:]         # the last layer's sumcheck has 2 L rounds: (b, c)
        w = [int(v) % R for v in circuit_evaluation[-1]]
        exponent = tau.powers_of_tau_in_g1.shape[0].bit_length() - 1
        blow = exponent - (len(w).bit_length() - 1)
        if blow < 0:
            raise ZkscError(-3, "the trusted setup is smaller than the input layer")
        poly = Multilinear(np.repeat(to_mont(w).reshape(-1, 4), 1 << blow, axis=0))
        b = list(ch[:L]) + [0] * (exponent - L)
        c = list(ch[L:]) + [0] * (exponent - L)
        commitment = MultilinearKZG.commitment(poly, tau, ctx)
        return commitment, SuccintGKRProof(base, MultilinearKZG.open(poly, b, tau, ctx), MultilinearKZG.open(poly, c, tau, ctx))

    @staticmethod
    def verify(circuit, commitment, proof, tau):   # gkr/src/succint_protocol.rs:169-266
        """the sumcheck chain of GKRProtocol::verify, then -- instead of evaluating the input layer -- the two KZG openings are checked
        (two pairing equations, host side: kzg.py / pairing.py) and their evaluations close the last claim"""
        from .kzg import MultilinearKZG
        if len(proof.sumcheck_proofs) != len(proof.wb_s) or len(proof.sumcheck_proofs) != len(proof.wc_s):
            return False
        transcript = FiatShamirTranscript()
        transcript.commit(b"".join(int(e).to_bytes(32, "big") for e in proof.w_0_mle.to_ints()))
        n_r = transcript.evaluate_n_challenge_into_field(proof.w_0_mle.n_vars)
        claimed = proof.w_0_mle.evaluation(n_r)
        p0 = proof.sumcheck_proofs[0]                     # generate_layer_one_verify_sumcheck, gkr/src/utils.rs:59-98
        if claimed != p0.sum:
            return False
        transcript.commit(p0.to_bytes())
        try:
            sub = MultiComposedSumcheckVerifier.verify_partial(p0)
        except ZkscError as e:
            if e.code == -7:
                return False
            raise
        add1, mul1 = circuit.add_mult_mle(0)
        rbc = list(n_r) + list(sub.challenges)
        wb, wc = proof.wb_s[0], proof.wc_s[0]
        if (add1.evaluation(rbc) * ((wb + wc) % R) + mul1.evaluation(rbc) * (wb * wc % R)) % R != sub.sum:
            return False
        alpha, beta = transcript.evaluate_challenge_into_field(), transcript.evaluate_challenge_into_field()
        claimed = (alpha * wb + beta * wc) % R
        r_b, r_c = [], []
        for i in range(1, len(proof.sumcheck_proofs)):    # :207-229 (alpha, beta are drawn before the sub-claim here)
            p = proof.sumcheck_proofs[i]
            if claimed != p.sum:
                return False
            transcript.commit(p.to_bytes())
            alpha, beta = transcript.evaluate_challenge_into_field(), transcript.evaluate_challenge_into_field()
            try:
                ch = MultiComposedSumcheckVerifier.verify_partial(p).challenges
            except ZkscError as e:
                if e.code == -7:
                    return False
                raise
            r_b, r_c = ch[:len(ch) // 2], ch[len(ch) // 2:]
            claimed = (alpha * proof.wb_s[i] + beta * proof.wc_s[i]) % R
        if proof.proof_wb_opening is None or proof.proof_wc_opening is None or commitment is None:
            return False
        n_tau = len(tau.powers_of_tau_in_g2)
        rb = list(r_b) + [0] * (n_tau - len(r_b))         # :231-241
        rc = list(r_c) + [0] * (n_tau - len(r_c))
        ok = MultilinearKZG.verify(commitment, rb, proof.proof_wb_opening, tau) and MultilinearKZG.verify(commitment, rc, proof.proof_wc_opening, tau)
        eb, ec = (from_mont(proof.proof_wb_opening.evaluation), from_mont(proof.proof_wc_opening.evaluation)) if ok else (0, 0)      # :248-256
        return claimed == (alpha * eb + beta * ec) % R
