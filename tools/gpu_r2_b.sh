#!/bin/bash
# round 2, session B: the resident rounds kernel (cooperative, from round >= 1): parity first, then the start-threshold sweep
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
for w in 1 60000000 120000000 250000000 500000000 1000000000 2000000000 4000000000; do
  for wl in c2 c3; do
    ZKSC_TAIL_WORK=$w timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/r2b_${wl}_w$w.json 2>gpurun_out/r2b_${wl}_w$w.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b_${wl}_w$w.json'))
    print('$wl tail_work=$w', 'ms/step %.4f'%d['ms_per_step'], 'launches', d['gpu_launches'], 'rounds us', [round(r['us'],1) for r in d['per_round']][:12])
except Exception as e:
    print('$wl $w FAILED', e, open('gpurun_out/r2b_${wl}_w$w.err').read()[-800:])
PY
  done
done
