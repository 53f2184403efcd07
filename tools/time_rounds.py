#!/usr/bin/env python
"""Per-round times (us, host-observed, zksc_ctx_round_times) of one proof of seeded tables: python tools/time_rounds.py n_vars degrees [proofs] [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_b200 as zk
n = int(sys.argv[1]); degs = [int(x) for x in sys.argv[2].split(",")]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
ctx = zk.Context(0)
t = zk.Tables.synth(ctx, n, degs, 1, n_proofs=B)
proto = zk.PROTO_SUMCHECK if degs == [1] else zk.PROTO_MULTI_PARTIAL
best = None
for it in range(reps + 2):
    t.reset()
    s = t.poly_sum()
    t.prove(proto, s)
    rt = np.array(ctx.round_times())
    if it >= 2:
        best = rt if best is None else np.minimum(best, rt)
print("n=%d degs=%s proofs=%d total %.1f us; rounds:" % (n, degs, B, best.sum()), [round(float(x), 1) for x in best[:12]])
