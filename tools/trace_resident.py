#!/usr/bin/env python
"""Device-side timeline of the resident rounds kernel (debug build: make -C zk_cryptography_b200/csrc OUT=... OBJDIR=... EXTRA=-DZKSC_RES_TRACE=1,
run with ZKSC_LIB=<that .so>): for every resident round of one proof, %globaltimer at the phase boundaries
  first CTA:     0 challenge in shared memory | 1 fold + evaluate done | 2 CTA-level reduction done | 3 counted (fence + atomic)
  last arriver:  4 detected it is last | 5 partials summed | 6 results published
printed as microseconds after phase 0 of the same round, plus the gap to the next round's phase 0 (= host turn-around + mailbox + relay).
usage: python tools/trace_resident.py [n_vars] [degree]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_b200 as zk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
degs = [int(sys.argv[2])] if len(sys.argv) > 2 else [2]
ctx = zk.Context(0)
t = zk.Tables.synth(ctx, n, degs, 1)
for it in range(3):
    t.reset()
    s = t.poly_sum()
    t.prove(zk.PROTO_MULTI_PARTIAL, s)
t.reset()
L = ctypes.CDLL(zk._lib.lib_path())
out = np.zeros(64 * 8, dtype=np.uint64)
rc = L.zksc_debug_res_trace(out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
assert rc == 0, rc
tr = out.reshape(64, 8).astype(np.int64)
print("round | ph1 fold+eval | ph2 cta-reduce | ph3 counted | ph4 last detected | ph5 summed | ph6 published | next round's ph0")
for r in range(64):
    if tr[r, 0] == 0:
        break
    base = tr[r, 0]
    row = [(tr[r, p] - base) / 1e3 if tr[r, p] >= base else float("nan") for p in range(1, 7)]
    nxt = (tr[r + 1, 0] - base) / 1e3 if r + 1 < 64 and tr[r + 1, 0] else float("nan")
    print("%2d | " % r + " ".join("%7.2f" % v for v in row) + " | %7.2f" % nxt)
