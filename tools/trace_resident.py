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

# per-CTA view of the first rounds: when did each CTA get the challenge, when were all its warps done, on which SM
if hasattr(L, "zksc_debug_res_cta_trace") and os.environ.get("ZKSC_TRACE_CTAS"):
    NC, NR = 640, 4
    o2 = np.zeros(NR * NC * 3, dtype=np.uint64)
    assert L.zksc_debug_res_cta_trace(o2.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))) == 0
    c = o2.reshape(NR, NC, 3).astype(np.int64)
    for r in range(NR):
        live = c[r, :, 0] > 0
        if not live.any():
            break
        t0 = c[r, live, 0].min()
        st = (c[r, live, 0] - t0) / 1e3
        en = (c[r, live, 1] - t0) / 1e3
        sm = c[r, live, 2]
        idx = np.nonzero(live)[0]
        print("round %d: %d CTAs; start after the first: median %.2f max %.2f us; done: min %.2f median %.2f p90 %.2f max %.2f us" % (
            r, live.sum(), np.median(st), st.max(), en.min(), np.median(en), np.percentile(en, 90), en.max()))
        order = np.argsort(en)
        print("   earliest CTAs (cta, sm, start, done):", [(int(idx[i]), int(sm[i]), round(float(st[i]), 1), round(float(en[i]), 1)) for i in order[:6]])
        print("   latest   CTAs (cta, sm, start, done):", [(int(idx[i]), int(sm[i]), round(float(st[i]), 1), round(float(en[i]), 1)) for i in order[-10:]])
        # by SM: mean finishing time of the SM's CTAs
        bysm = {}
        for i in range(len(en)):
            bysm.setdefault(int(sm[i]), []).append(float(en[i]))
        m = sorted((np.mean(v), k, len(v)) for k, v in bysm.items())
        print("   SMs by mean done time: fastest", [(k, round(a, 1), n) for a, k, n in m[:5]], "slowest", [(k, round(a, 1), n) for a, k, n in m[-8:]])
        dur = en - st
        print("   duration (done - start): min %.2f median %.2f max %.2f; CTAs with the extra pair (ci < extra): see index order" % (dur.min(), np.median(dur), dur.max()))
        print("   mean done by CTA-index decile:", [round(float(np.mean(en[(idx >= lo) & (idx < lo + 64)])), 1) for lo in range(0, int(idx.max()) + 1, 64)])
