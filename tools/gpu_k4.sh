#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for e in "ZKSC_ROUND_STATIC=1" "ZKSC_ROUND_STATIC=0"; do
  echo "== $e"
  for a in "27 1" "26 1" "24 1" "25 2" "22 2,2" "20 2,3" "24 4" "22 2 4"; do env $e python tools/time_rounds.py $a 2>&1 | tail -1; done
done
ZKSC_AB_WORKLOADS="c5 c4 c1" bash tools/gpu_ab_env.sh "rstatic:ZKSC_ROUND_STATIC=1" "rdyn:ZKSC_ROUND_STATIC=0"
ZKSC_AB_WORKLOADS="c2 c3" bash tools/gpu_ab_env.sh "w600:ZKSC_TAIL_WORK=600000000" "w1100:ZKSC_TAIL_WORK=1100000000" "w2200:ZKSC_TAIL_WORK=2200000000"
