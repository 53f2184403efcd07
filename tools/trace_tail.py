#!/usr/bin/env python
"""Device-side timeline of the resident rounds kernel (debug build: make ... EXTRA=-DZKSC_TAIL_TRACE=1, ZKSC_LIB=<that .so>):
for every resident round of one c2-shaped proof, the %globaltimer stamps of CTA 0 and CTA 1 at the phase boundaries
  0 round entered | 1 challenge in shared memory | 2 acquire fence done | 3 fold + evaluate done | 4 release fence done
  5 CTA sums handed over | 6 CTA 0 has every CTA's sums | 7 results published
printed as microseconds relative to CTA 0's phase 0 of the same round.  usage: python tools/trace_tail.py [n_vars] [degrees]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_b200 as zk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
degs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2]
ctx = zk.Context(0)
t = zk.Tables.synth(ctx, n, degs, 1)
for it in range(3):
    t.reset()
    s = t.poly_sum()
    t.prove(zk.PROTO_MULTI_PARTIAL, s)
t.reset()
L = ctypes.CDLL(zk._lib.lib_path())
out = np.zeros(64 * 2 * 8, dtype=np.uint64)
rc = L.zksc_debug_tail_trace(out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
assert rc == 0, rc
tr = out.reshape(64, 2, 8).astype(np.int64)
print("resident round | pairs | CTA0 phases 1..7 (us after its phase 0) | CTA1 phases 0..5 (us after CTA0's phase 0) | gap to next round's phase 0")
first_pairs = None
for r in range(64):
    if tr[r, 0, 0] == 0:
        break
    base = tr[r, 0, 0]
    c0 = [(tr[r, 0, p] - base) / 1e3 if tr[r, 0, p] >= base else float("nan") for p in range(1, 8)]
    c1 = [(tr[r, 1, p] - base) / 1e3 if tr[r, 1, p] >= base else float("nan") for p in range(0, 6)]
    nxt = (tr[r + 1, 0, 0] - base) / 1e3 if r + 1 < 64 and tr[r + 1, 0, 0] else float("nan")
    print("%2d | " % r + " ".join("%6.2f" % v for v in c0) + " | " + " ".join("%6.2f" % v for v in c1) + " | %6.2f" % nxt)
