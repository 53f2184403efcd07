#!/bin/bash
# parity tests + the short bench lines (no CPU leg, no target leg); usage: bash tools/gpu_quick.sh [label:ENV=1 ...]
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_gkr.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -3
for spec in "${@:-default:ZKSC_X=0}"; do
  label=${spec%%:*}; envs=${spec#*:}
  ZKSC_AB_WORKLOADS="c2 c1 c5" bash tools/gpu_ab_env.sh "$spec"
  for wl in c4 c4b; do env $envs timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu > gpurun_out/quick_${wl}_$label.json 2>gpurun_out/quick_${wl}_$label.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/quick_${wl}_$label.json') if l.startswith('{')][-1]); print('$label $wl ms/step %.4f' % d['ms_per_step'], d['latency']['us_per_round_incl_layer_setup'])"; done
done
