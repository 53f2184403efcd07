#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_ab_env.sh "static:ZKSC_RES_STATIC=1" "dynamic:ZKSC_RES_STATIC=0" "dyn_early:ZKSC_TAIL_WORK=4000000000" "static_early:ZKSC_RES_STATIC=1 ZKSC_TAIL_WORK=4000000000"
