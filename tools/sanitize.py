#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family once, small sizes.
usage: compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_b200 as zk
from zk_cryptography_b200._lib import proof_to_bytes

ctx = zk.Context(0)
zk.set_default_context(ctx)
for n, degs, B in ((13, [2], 1), (12, [3], 1), (12, [2, 2], 2), (11, [1], 1), (10, [2, 3], 3), (12, [4], 1)):
    t = zk.Tables.synth(ctx, n, degs, 7, n_proofs=B)
    s = t.poly_sum()
    proto = zk.PROTO_SUMCHECK if degs == [1] else zk.PROTO_MULTI_PARTIAL
    msgs, lens, chal = t.prove(proto, s)
    print("prove", n, degs, B, len(proof_to_bytes(proto, msgs[0], lens[0])), "bytes")
    t.free()
lc = zk.LayeredCircuit.random([3, 6, 6, 7], 3)
inp = list(range(1, 129))
lc.evaluate(inp)
p = lc.prove()
print("layered GKR verifies:", lc.verify(inp, p))
c = zk.Circuit.random(5)
ev = c.evaluation(list(range(1, 33)))
print("dense GKR verifies:", zk.GKRProtocol.verify(c, list(range(1, 33)), zk.GKRProtocol.prove(c, ev)))
print("SANITIZE RUN DONE")
