"""CPU: the GKR / circuit oracle model (oracle/gkrmodel.py) against the reference's own known answers
(circuit/src/circuit.rs:139-518, circuit/src/utils.rs:38-64, gkr/src/protocol.rs:209-285) and against itself
(literal dense restatement == closed-form wiring tables, incl. layer sumchecks through the C oracle)."""
import pytest

from oracle import gkrmodel as g
from oracle import pymodel as pm

R = pm.R_MOD
G, CL, C = g.Gate, g.CircuitLayer, g.Circuit


def circuit_1():   # gkr/src/protocol.rs:210-224
    return C([CL([G(g.MUL, [0, 1])]), CL([G(g.ADD, [0, 1]), G(g.MUL, [2, 3])])]), [2, 3, 4, 5]


def circuit_2():   # gkr/src/protocol.rs:236-275
    c = C([CL([G(g.ADD, [0, 1])]), CL([G(g.MUL, [0, 1]), G(g.ADD, [2, 3])]),
           CL([G(g.ADD, [0, 1]), G(g.MUL, [2, 3]), G(g.MUL, [4, 5]), G(g.MUL, [6, 7])]),
           CL([G(g.MUL, [0, 1]), G(g.MUL, [2, 3]), G(g.MUL, [4, 5]), G(g.ADD, [6, 7]), G(g.MUL, [8, 9]), G(g.ADD, [10, 11]), G(g.MUL, [12, 13]),
               G(g.MUL, [14, 15])])])
    return c, [2, 1, 3, 1, 4, 1, 2, 2, 3, 3, 4, 4, 2, 3, 3, 4]


def test_circuit_evaluation_kats():   # circuit/src/circuit.rs:139-243
    c, inp = circuit_1()
    assert c.evaluation(inp) == [[100], [5, 20], [2, 3, 4, 5]]
    c = C([CL([G(g.MUL, [0, 1]), G(g.MUL, [2, 3])]), CL([G(g.MUL, [0, 0]), G(g.MUL, [1, 1]), G(g.MUL, [1, 2]), G(g.MUL, [3, 3])])])
    assert c.evaluation([3, 2, 3, 1]) == [[36, 6], [9, 4, 6, 1], [3, 2, 3, 1]]
    c = C([CL([G(g.ADD, [0, 1])]), CL([G(g.ADD, [0, 1]), G(g.MUL, [2, 3])]), CL([G(g.ADD, [0, 1]), G(g.MUL, [2, 3]), G(g.MUL, [4, 5]), G(g.MUL, [6, 7])])])
    assert c.evaluation([2, 3, 1, 4, 1, 2, 3, 4]) == [[33], [9, 24], [5, 4, 2, 12], [2, 3, 1, 4, 1, 2, 3, 4]]
    c2, inp2 = circuit_2()
    assert c2.evaluation(inp2)[0][0] == 224


def test_wiring_table_kats():   # circuit/src/utils.rs:38-64, circuit/src/circuit.rs:246-330
    assert [g.size_of_mle_n_var_at_each_layer(i) for i in range(5)] == [8, 32, 256, 2048, 16384]
    assert g.transform_label_to_binary_and_to_decimal(1, 1, 2, 3) == 27
    assert g.transform_label_to_binary_and_to_decimal(2, 1, 2, 3) == 83
    assert [g.binary_string(0, 0), g.binary_string(0, 1), g.binary_string(0, 2), g.binary_string(5, 3)] == ["0", "0", "00", "101"]
    c = C([CL([G(g.ADD, [0, 1])]), CL([G(g.ADD, [0, 1]), G(g.MUL, [2, 3])]), CL([G(g.ADD, [0, 1]), G(g.MUL, [2, 3]), G(g.MUL, [4, 5]), G(g.MUL, [6, 7])])])
    add, mul = c.add_mult_mle(0)
    assert sum(mul.evaluations) == 0 and sum(add.evaluations) == 1
    assert add.evaluation([0, 0, 1]) == 1
    for pt in ([0, 0, 0], [1, 0, 0], [1, 0, 1], [1, 1, 1]):
        assert add.evaluation(pt) == 0


@pytest.mark.parametrize("make", [circuit_1, circuit_2, lambda: (C.random(5), list(range(1, 33)))])
def test_literal_and_closed_form_provers_agree(make):
    c, inp = make()
    ev = c.evaluation(inp)
    lit = g.GKRProtocol.prove(c, ev)
    assert g.GKRProtocol.verify(c, inp, lit)
    assert g.GKRProtocol.prove_sparse(c, ev).to_bytes() == lit.to_bytes()
    lit.wc_s[-1] = (lit.wc_s[-1] + 1) % R
    assert not g.GKRProtocol.verify(c, inp, lit)


def test_closed_form_prover_through_the_c_oracle():
    c = C.random(6)
    inp = [(7 * i + 3) % 1000 for i in range(64)]
    ev = c.evaluation(inp)
    a = g.GKRProtocol.prove_sparse(c, ev)
    b = g.GKRProtocol.prove_sparse(c, ev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate)
    assert a.to_bytes() == b.to_bytes()


# ---- layered circuits of any widths (oracle side of zksc_gkr_prove_linear) -------------------------------------------------------
def _random_layered(log_width, seed):
    import random
    rng = random.Random(seed)
    return g.LayeredCircuit(log_width, [[(rng.randrange(2), rng.randrange(1 << log_width[i + 1]), rng.randrange(1 << log_width[i + 1]))
                                         for _ in range(1 << log_width[i])] for i in range(len(log_width) - 1)])


def test_layered_prover_is_the_reference_prover_on_pyramids():
    """prove_layered (label widths taken from the layers) == the literal restatement of GKRProtocol::prove on the reference's own
    circuits and on Circuit::random -- the anchor for the widths the reference cannot hold"""
    for c, inp in (circuit_1(), circuit_2(), (g.Circuit.random(4), [(7 * i + 3) % R for i in range(16)])):
        ev = c.evaluation(inp)
        lc = g.LayeredCircuit.from_circuit(c)
        assert lc.evaluation(inp) == ev
        assert g.prove_layered(lc, ev).to_bytes() == g.GKRProtocol.prove(c, ev).to_bytes()


def test_layered_prover_python_and_c_layer_provers_agree():
    lc = _random_layered([2, 3, 3, 4], 5)
    inp = [(0x9E3779B97F4A7C15 * (i + 1)) % R for i in range(16)]
    ev = lc.evaluation(inp)
    assert [len(l) for l in ev] == [4, 8, 8, 16]
    a = g.prove_layered(lc, ev)
    b = g.prove_layered(lc, ev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate)
    assert a.to_bytes() == b.to_bytes() and a.wb_s == b.wb_s and a.wc_s == b.wc_s
