#!/bin/bash
cd "$GRAFT_REPO_ROOT"
ZKSC_PROFILE=1 timeout 300 python - <<'PY' 2>&1 | grep "gkr linear" | tail -12
import time
import zk_cryptography_b200 as zk
lc = zk.LayeredCircuit.random([20]*4, 1)
lc.evaluate([i+1 for i in range(1<<20)])
lc.prove_raw(); 
t0=time.perf_counter(); lc.prove_raw(); print("gkr linear wall", (time.perf_counter()-t0)*1e3, "ms")
PY
