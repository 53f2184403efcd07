"""Debug: staged (TMA) vs plain round kernels, round by round, fixed challenges."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_b200 as zk

os.environ["ZKSC_NO_STAGED"] = "1"
plain = zk.Context(0)
os.environ["ZKSC_NO_STAGED"] = "0"
os.environ["ZKSC_STAGED_FOLD"] = sys.argv[1] if len(sys.argv) > 1 else "1"
staged = zk.Context(0)
rng = np.random.default_rng(5)
bad = 0
for n, degs in [(13, [1]), (14, [1]), (17, [1]), (18, [1]), (19, [1]), (20, [1]), (18, [2]), (20, [2]), (19, [3]), (18, [2, 3])]:
    ta = zk.Tables.synth(plain, n, degs, 77)
    tb = zk.Tables.synth(staged, n, degs, 77)
    for rnd in range(n):
        ea, eb = ta.round_evals(), tb.round_evals()
        if not np.array_equal(ea, eb):
            print("MISMATCH n=%d degs=%s round=%d (table size 2^%d)" % (n, degs, rnd, n - rnd), flush=True)
            bad += 1
            break
        c = zk.to_mont([int(rng.integers(1, 2**62)) * 0x9E3779B97F4A7C15 % zk.R_MOD])
        ta.bind(c); tb.bind(c)
    else:
        ra, rb = ta.residual(), tb.residual()
        ok = np.array_equal(ra, rb)
        print("n=%d degs=%s: %s" % (n, degs, "ok" if ok else "RESIDUAL MISMATCH"), flush=True)
        bad += 0 if ok else 1
    ta.free(); tb.free()
print("bad =", bad)
