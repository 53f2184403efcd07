"""GPU: the GKR layer driver (zk_cryptography_b200/gkr.py: tables built on the device, layer sumchecks in the CUDA round
kernels) against the oracle model -- proof bytes, claimed evaluations, verifier decisions.  Mirrors
gkr/src/protocol.rs:209-285 (the reference only asserts verify == true; here every byte is compared as well)."""
import pytest

import zk_cryptography_b200 as zk
from oracle import gkrmodel as g

pytestmark = pytest.mark.gpu
R = zk.R_MOD


@pytest.fixture(scope="module", autouse=True)
def _ctx(ctx):
    return ctx


def both(layers):
    """the same circuit for the product API and the oracle model"""
    zc = zk.Circuit([zk.CircuitLayer([zk.Gate(zk.GateType.Add if t == "Add" else zk.GateType.Mul, i) for t, i in layer]) for layer in layers])
    oc = g.Circuit([g.CircuitLayer([g.Gate(t, i) for t, i in layer]) for layer in layers])
    return zc, oc


CIRCUIT_1 = [[("Mul", [0, 1])], [("Add", [0, 1]), ("Mul", [2, 3])]]                       # protocol.rs:210-224
CIRCUIT_2 = [[("Add", [0, 1])], [("Mul", [0, 1]), ("Add", [2, 3])],                       # protocol.rs:236-275
             [("Add", [0, 1]), ("Mul", [2, 3]), ("Mul", [4, 5]), ("Mul", [6, 7])],
             [("Mul", [0, 1]), ("Mul", [2, 3]), ("Mul", [4, 5]), ("Add", [6, 7]), ("Mul", [8, 9]), ("Add", [10, 11]), ("Mul", [12, 13]), ("Mul", [14, 15])]]


@pytest.mark.parametrize("layers,inp", [(CIRCUIT_1, [2, 3, 4, 5]), (CIRCUIT_2, [2, 1, 3, 1, 4, 1, 2, 2, 3, 3, 4, 4, 2, 3, 3, 4])])
def test_gkr_protocol_reference_cases(layers, inp):
    zc, oc = both(layers)
    ev = zc.evaluation(inp)
    assert ev == oc.evaluation(inp)
    proof = zk.GKRProtocol.prove(zc, ev)
    want = g.GKRProtocol.prove(oc, ev)
    assert proof.to_bytes() == want.to_bytes()
    assert proof.wb_s == want.wb_s and proof.wc_s == want.wc_s
    assert zk.GKRProtocol.verify(zc, inp, proof)
    proof.wb_s[-1] = (proof.wb_s[-1] + 1) % R
    assert not zk.GKRProtocol.verify(zc, inp, proof)


def test_wiring_tables_and_circuit_api():
    zc, oc = both(CIRCUIT_2)
    for li in range(3):
        za, zm = zc.add_mult_mle(li)
        oa, om = oc.add_mult_mle(li)
        assert za.to_ints() == oa.evaluations and zm.to_ints() == om.evaluations
    r5 = zk.Circuit.random(5)
    o5 = g.Circuit.random(5)
    assert [[(gt.gate_type, gt.inputs) for gt in l.layer] for l in r5.layers] == [[(gt.gate_type, gt.inputs) for gt in l.layer] for l in o5.layers]
    with pytest.raises(zk.ZkscError):
        zk.GKRProtocol.prove(zk.Circuit([zk.CircuitLayer([zk.Gate(zk.GateType.Add, [0, 5])])]), [[1], [1, 2]])   # input label out of range


@pytest.mark.parametrize("depth", [5, 8])
def test_gkr_random_circuit_vs_oracle(depth):
    """Circuit::random(depth) (the reference's bench circuit, gkr/benches/gkr_benchmark.rs:11-20) on pseudo-random inputs; the
    oracle's layer sumchecks run in the C oracle so that depth 8 (2^16-entry layer tables) stays fast."""
    zc, oc = zk.Circuit.random(depth), g.Circuit.random(depth)
    inp = [(0x9E3779B97F4A7C15 * (i + 1)) % R for i in range(1 << depth)]
    ev = zc.evaluation(inp)
    proof = zk.GKRProtocol.prove(zc, ev)
    want = g.GKRProtocol.prove_sparse(oc, ev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate)
    assert proof.to_bytes() == want.to_bytes()
    assert zk.GKRProtocol.verify(zc, inp, proof)


def test_gkr_full_width_roundtrip():
    """BASELINE config 4 at the reference's circuit shape: 10 layers, 2^10 inputs, largest layer sumcheck 2^20 entries.
    prove -> verify closes; a second prove gives the same bytes (size-independent properties; the oracle is compared at depth <= 8)."""
    zc = zk.Circuit.random(10)
    inp = [(0xD1342543DE82EF95 * (i + 7)) % R for i in range(1 << 10)]
    ev = zc.evaluation(inp)
    proof = zk.GKRProtocol.prove(zc, ev)
    assert zk.GKRProtocol.verify(zc, inp, proof)
    assert zk.GKRProtocol.prove(zc, ev).to_bytes() == proof.to_bytes()
    assert len(proof.sumcheck_proofs) == 10 and len(proof.sumcheck_proofs[-1].round_polys) == 20


@pytest.mark.parametrize("depth", [1, 2, 3, 6])
def test_gkr_c_driver_matches_layerwise_driver(depth):
    """zksc_gkr_prove (the whole proof in one C call) against the same protocol driven layer by layer over the finer-grained
    entry points: same bytes, same claimed evaluations, same sums."""
    zc = zk.Circuit.random(depth)
    inp = [(0xA24BAED4963EE407 * (i + 3)) % R for i in range(1 << depth)]
    ev = zc.evaluation(inp)
    a, b = zk.GKRProtocol.prove(zc, ev), zk.GKRProtocol.prove_layerwise(zc, ev)
    assert a.to_bytes() == b.to_bytes()
    assert a.wb_s == b.wb_s and a.wc_s == b.wc_s
    assert [p.sum for p in a.sumcheck_proofs] == [p.sum for p in b.sumcheck_proofs]
    if depth >= 2:   # with one layer the reference's verifier never sets r_b / r_c and panics in evaluation (protocol.rs:133-186)
        assert zk.GKRProtocol.verify(zc, inp, a)
    else:
        with pytest.raises(zk.ZkscError):
            zk.GKRProtocol.verify(zc, inp, a)


@pytest.mark.parametrize("linear_min", ["1", "3", "99"])
def test_gkr_c_driver_two_phase_form_of_large_layers(built, linear_min):
    """zksc_gkr_prove runs its large layers (default: 9 bits per input label and up) in the two-phase form of csrc/gkr_linear.cuh: with
    every layer / the layers from 3 bits / no layer in that form the bytes must be the oracle's"""
    import os
    old = os.environ.get("ZKSC_GKR_LINEAR_MIN")
    os.environ["ZKSC_GKR_LINEAR_MIN"] = linear_min
    try:
        c = zk.Context(0)
        try:
            for layers, inp in [(CIRCUIT_1, [2, 3, 4, 5]), (CIRCUIT_2, [2, 1, 3, 1, 4, 1, 2, 2, 3, 3, 4, 4, 2, 3, 3, 4])]:
                zc, oc = both(layers)
                ev = zc.evaluation(inp)
                got, want = zk.GKRProtocol.prove(zc, ev, ctx=c), g.GKRProtocol.prove(oc, ev)
                assert got.to_bytes() == want.to_bytes() and got.wb_s == want.wb_s and got.wc_s == want.wc_s
            zc, oc = zk.Circuit.random(7), g.Circuit.random(7)
            inp = [(0x9E3779B97F4A7C15 * (i + 1)) % R for i in range(1 << 7)]
            ev = zc.evaluation(inp)
            got = zk.GKRProtocol.prove(zc, ev, ctx=c)
            assert got.to_bytes() == g.GKRProtocol.prove_sparse(oc, ev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate).to_bytes()
        finally:
            c.close()
    finally:
        if old is None:
            os.environ.pop("ZKSC_GKR_LINEAR_MIN", None)
        else:
            os.environ["ZKSC_GKR_LINEAR_MIN"] = old


def test_gkr_c_driver_reference_cases_layerwise_too():
    for layers, inp in [(CIRCUIT_1, [2, 3, 4, 5]), (CIRCUIT_2, [2, 1, 3, 1, 4, 1, 2, 2, 3, 3, 4, 4, 2, 3, 3, 4])]:
        zc, oc = both(layers)
        ev = zc.evaluation(inp)
        assert zk.GKRProtocol.prove_layerwise(zc, ev).to_bytes() == g.GKRProtocol.prove(oc, ev).to_bytes()


def test_gkr_c_driver_shape_errors():
    """what the reference's constructors panic on (Multilinear::new power-of-two, ComposedMultilinear::new equal arity,
    label widths of circuit/src/utils.rs:12-34) comes back as ZKSC_ERR_SHAPE"""
    zc = zk.Circuit.random(3)
    ev = zc.evaluation(list(range(1, 9)))
    bad = [list(l) for l in ev]
    bad[2] = bad[2][:3]                                   # not a power of two
    with pytest.raises(zk.ZkscError) as e:
        zk.GKRProtocol.prove(zc, bad)
    assert e.value.code == -3
    with pytest.raises(zk.ZkscError):
        zk.GKRProtocol.prove(zc, ev[:-1])                 # a layer is missing
    two_out = zk.Circuit([zk.CircuitLayer([zk.Gate(zk.GateType.Add, [0, 1]), zk.Gate(zk.GateType.Mul, [0, 1])])])
    with pytest.raises(zk.ZkscError):
        zk.GKRProtocol.prove(two_out, [[3, 2], [1, 2]])   # two output gates: w_0 would have 3 entries


# ---- layered circuits of any widths, linear-time prover (zksc_gkr_prove_linear; BASELINE config 4 as written) ---------------------
def _olayered(lc):
    return g.LayeredCircuit(lc.log_width, [[(int(t), int(a), int(b)) for t, a, b in zip(*layer)] for layer in lc.layers()])


@pytest.mark.parametrize("layers,inp", [(CIRCUIT_1, [2, 3, 4, 5]), (CIRCUIT_2, [2, 1, 3, 1, 4, 1, 2, 2, 3, 3, 4, 4, 2, 3, 3, 4])])
def test_linear_prover_reference_cases(layers, inp):
    """on the reference's own circuits the linear-time prover gives the bytes of GKRProtocol::prove (oracle, literal restatement) and of
    the dense device prover, and GKRProtocol::verify accepts them (gkr/src/protocol.rs:209-285)"""
    zc, oc = both(layers)
    lc = zk.LayeredCircuit.from_circuit(zc)
    ev = oc.evaluation(inp)
    assert lc.evaluate(inp) == ev[0]
    assert [lc.layer_values(i) for i in range(len(ev))] == ev
    proof = lc.prove()
    want = g.GKRProtocol.prove(oc, ev)
    assert proof.to_bytes() == want.to_bytes()
    assert proof.wb_s == want.wb_s and proof.wc_s == want.wc_s
    assert proof.to_bytes() == zk.GKRProtocol.prove(zc, ev).to_bytes()
    assert zk.GKRProtocol.verify(zc, inp, proof) and lc.verify(inp, proof)
    proof.wb_s[-1] = (proof.wb_s[-1] + 1) % R
    assert not lc.verify(inp, proof)


@pytest.mark.parametrize("depth", [5, 9])
def test_linear_prover_random_pyramid_vs_dense_device_prover(depth):
    """Circuit::random(depth): linear-time prover == dense device prover == oracle (C layer sumchecks), byte for byte"""
    zc, oc = zk.Circuit.random(depth), g.Circuit.random(depth)
    inp = [(0x9E3779B97F4A7C15 * (i + 1)) % R for i in range(1 << depth)]
    lc = zk.LayeredCircuit.from_circuit(zc)
    lc.evaluate(inp)
    ev = zc.evaluation(inp)
    proof = lc.prove()
    assert proof.to_bytes() == zk.GKRProtocol.prove(zc, ev).to_bytes()
    if depth <= 8:
        assert proof.to_bytes() == g.GKRProtocol.prove_sparse(oc, ev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate).to_bytes()
    assert zk.GKRProtocol.verify(zc, inp, proof)


@pytest.mark.parametrize("log_width,seed", [([3, 3, 3], 1), ([0, 4, 4, 4], 2), ([2, 5, 3, 6], 3), ([6, 6, 6, 6, 6], 4), ([4, 8, 8], 5), ([1, 1, 1], 6)])
def test_linear_prover_uniform_and_ragged_widths_vs_dense_oracle(log_width, seed):
    """widths the reference's Circuit cannot hold: against the oracle's DENSE prover (width^2-entry layer tables through the C oracle)"""
    lc = zk.LayeredCircuit.random(log_width, seed)
    oc = _olayered(lc)
    inp = [(0xD1B54A32D192ED03 * (i + 1) + seed) % R for i in range(1 << log_width[-1])]
    ev = oc.evaluation(inp)
    assert lc.evaluate(inp) == ev[0]
    proof = lc.prove()
    want = g.prove_layered(oc, ev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate)
    assert proof.to_bytes() == want.to_bytes()
    assert proof.wb_s == want.wb_s and proof.wc_s == want.wc_s
    assert lc.verify(inp, proof) and lc.verify(inp, proof, device=False)       # wiring polynomials on the device / in Python integers
    ch = proof.challenges[:2 * log_width[1]]
    pts = [(ch[:log_width[1]][:max(log_width[0], 1)] if False else [5] * max(log_width[0], 1), 3), ([7] * max(log_width[0], 1), 11)]
    b, c = ch[:log_width[1]], ch[log_width[1]:]
    assert lc.wiring_at_device(0, pts, b, c) == lc.wiring_at(0, pts, b, c)
    assert lc.wiring_at_device(0, pts[:1], b, c) == lc.wiring_at(0, pts[:1], b, c)
    proof.sumcheck_proofs[0].round_polys[0].monomial[0] = ((proof.sumcheck_proofs[0].round_polys[0].monomial[0][0] + 1) % R, proof.sumcheck_proofs[0].round_polys[0].monomial[0][1])
    assert not lc.verify(inp, proof)


def test_linear_prover_special_inputs_and_repeated_proofs():
    """all-zero and all-one inputs (every round polynomial drops monomials), one circuit proved twice with different inputs"""
    lc = zk.LayeredCircuit.random([3, 4, 4], 9)
    oc = _olayered(lc)
    for inp in ([0] * 16, [1] * 16, list(range(16))):
        ev = oc.evaluation(inp)
        lc.evaluate(inp)
        proof = lc.prove()
        assert proof.to_bytes() == g.prove_layered(oc, ev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate).to_bytes()
        assert lc.verify(inp, proof)


def test_linear_prover_width_2_16_verifies():
    """width 2^16 x 3 layers (the dense form would need 2^32-entry tables): the proof must pass the verifier"""
    lc = zk.LayeredCircuit.random([16, 16, 16, 16], 11)
    inp = [(0x9E3779B97F4A7C15 * (i + 1)) % R for i in range(1 << 16)]
    lc.evaluate(inp)
    proof = lc.prove()
    assert lc.verify(inp, proof) and lc.verify(inp, proof, device=False)
    proof.wc_s[1] = (proof.wc_s[1] + 1) % R
    assert not lc.verify(inp, proof)


def test_linear_prover_shape_errors():
    with pytest.raises(zk.ZkscError):
        zk.LayeredCircuit([1, 1], [0, 0], [0, 2], [0, 0])          # input label out of range
    with pytest.raises(zk.ZkscError):
        zk.LayeredCircuit([1, 1], [0, 2], [0, 1], [0, 0])          # unknown gate type
    with pytest.raises(zk.ZkscError):
        zk.LayeredCircuit([1, 1], [0], [0], [0])                   # 2^1 gates announced, one given
    lc = zk.LayeredCircuit([1, 1], [0, 1], [0, 1], [1, 0])
    with pytest.raises(zk.ZkscError):
        lc.prove()                                                 # nothing evaluated yet
    with pytest.raises(zk.ZkscError):
        lc.evaluate([1, 2, 3])                                     # wrong input length
