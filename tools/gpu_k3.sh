#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_gkr.py -m gpu -x -q 2>&1 | tail -3
ZKSC_AB_WORKLOADS="c2 c3 c5" bash tools/gpu_ab_env.sh "rstatic:ZKSC_ROUND_STATIC=1" "rdyn:ZKSC_ROUND_STATIC=0" "rstatic2:ZKSC_ROUND_STATIC=1" "rdyn2:ZKSC_ROUND_STATIC=0"
