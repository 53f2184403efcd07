#!/bin/bash
# round 2, session C: resident kernel v2 (direct polling, fence-free relay): parity + c1/c2/c3/c4/c5 numbers
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
for wl in c2 c3 c1 c5 c4; do
    timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/r2c_${wl}.json 2>gpurun_out/r2c_${wl}.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c_${wl}.json'))
    print('$wl', 'ms/step %.4f'%d['ms_per_step'], 'launches', d['gpu_launches'], 'rounds us', [round(r['us'],1) for r in (d.get('per_round') or [])])
except Exception as e:
    print('$wl FAILED', e, open('gpurun_out/r2c_${wl}.err').read()[-800:])
PY
done
for w in 250000000 500000000 2000000000; do
    ZKSC_TAIL_WORK=$w timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/r2c_c2_w$w.json 2>/dev/null
    python -c "
import json
d=json.load(open('gpurun_out/r2c_c2_w$w.json')); print('c2 tail_work=$w ms/step %.4f'%d['ms_per_step'], [round(r['us'],1) for r in d['per_round']][:10])"
done
