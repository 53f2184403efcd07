#!/bin/bash
cd "$GRAFT_REPO_ROOT"
ZKSC_AB_WORKLOADS="c5 c4" bash tools/gpu_ab_env.sh "g8:ZKSC_DYN_MAX_GROUPS=8" "g64:ZKSC_DYN_MAX_GROUPS=64" "g0:ZKSC_DYN_MAX_GROUPS=0"
for e in ZKSC_DYN_MAX_GROUPS=8 ZKSC_DYN_MAX_GROUPS=64; do for a in "22 2 4" "22 2 16" "20 2,2 8" "24 3 2"; do env $e python tools/time_rounds.py $a 2>&1 | tail -1; done; done
