#!/bin/bash
# GKR tests + the c4b line (tools/gpu_suite.sh runs both as part of everything)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_gkr.py -m gpu -x -q 2>&1 | tail -12
timeout 900 python bench.py --workload c4b --steps 10 --warmup 3 > gpurun_out/gkr_bench_c4b.json 2> gpurun_out/gkr_bench_c4b.err; echo "c4b rc=$?"; tail -5 gpurun_out/gkr_bench_c4b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/gkr_bench_c4b.json') if l.startswith('{')][-1])
print('c4b value %.3e ms %.4f e2e %s launches %s' % (d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches']))
print(d['parity'], d['verified'])
PY
