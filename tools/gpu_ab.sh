#!/bin/bash
# A/B of library variants (ZKSC_LIB): c2 and c3 on one B200, per-round times; usage: bash tools/gpu_ab.sh default k3 k2 ...
cd "$GRAFT_REPO_ROOT"
for v in "$@"; do
  if [ $v = default ]; then unset ZKSC_LIB; else export ZKSC_LIB=$PWD/build/libzksc_$v.so; fi
  for wl in c2 c3; do
    timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/ab_${wl}_$v.json 2>gpurun_out/ab_${wl}_$v.err
    python -c "
import json
d=json.load(open('gpurun_out/ab_${wl}_$v.json')); print('$v $wl ms/step %.4f sha %s'%(d['ms_per_step'], d['proof_sha256'][:12]), [round(r['us'],1) for r in d['per_round']][:9])" || tail -3 gpurun_out/ab_${wl}_$v.err
  done
done
