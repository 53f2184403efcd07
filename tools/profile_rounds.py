"""Host-side wall-clock split of every round of one proof (ZKSC_PROFILE=1): launch / wait / transcript / bind."""
import os, sys
os.environ["ZKSC_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_b200 as zk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
degs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2]
ctx = zk.Context(0)
t = zk.Tables.synth(ctx, n, degs, 1)
for it in range(3):
    t.reset()
    s = t.poly_sum()
    if it == 2:
        sys.stderr.write("---- measured proof ----\n")
    t.prove(zk.PROTO_MULTI_PARTIAL, s)
