#!/bin/bash
# resident-kernel variants for degree 3: 3 CTAs per SM (no spills) and/or the single-conditional-subtraction fold
cd "$GRAFT_REPO_ROOT"
for v in default b3 s3 b3s3; do
  if [ $v = default ]; then unset ZKSC_LIB; else export ZKSC_LIB=$PWD/build/libzksc_$v.so; fi
  for w in 1000000000 8000000000; do
    ZKSC_TAIL_WORK=$w timeout 300 python bench.py --workload c3 --steps 6 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/r2v_c3_${v}_$w.json 2>/dev/null
    python -c "
import json
d=json.load(open('gpurun_out/r2v_c3_${v}_$w.json')); print('$v tail_work=$w c3 ms/step %.3f'%d['ms_per_step'], [round(r['us'],1) for r in d['per_round']][:14])"
  done
done
