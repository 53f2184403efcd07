#!/bin/bash
# A/B of run-time switches on one B200: per-round times of c2, c3 and c5; usage: bash tools/gpu_ab_env.sh "label:VAR=1 VAR2=x" ...
cd "$GRAFT_REPO_ROOT"
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  for wl in ${ZKSC_AB_WORKLOADS:-c2 c3}; do
    env $envs timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/ab_${wl}_$label.json 2>gpurun_out/ab_${wl}_$label.err
    python - "$label" "$wl" <<'PY' || tail -3 gpurun_out/ab_${wl}_$label.err
import json, sys
label, wl = sys.argv[1:3]
d = json.loads([l for l in open('gpurun_out/ab_%s_%s.json' % (wl, label)) if l.startswith('{')][-1])
print('%-12s %s ms/step %.4f sha %s' % (label, wl, d['ms_per_step'], d['proof_sha256'][:12]), [round(r['us'], 1) for r in (d.get('per_round') or [])][:16])
PY
  done
done
