#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_gkr.py -m gpu -x -q 2>&1 | tail -3
for a in "27 1" "24 1" "25 2" "22 2,2" "20 2,3" "24 4" "22 2 4"; do python tools/time_rounds.py $a 2>&1 | tail -1; done
ZKSC_AB_WORKLOADS="c2 c3 c5 c1" bash tools/gpu_ab_env.sh "allstatic:ZKSC_ROUND_STATIC=1 ZKSC_RES_STATIC=1" "dyn4:ZKSC_ROUND_STATIC=0"
