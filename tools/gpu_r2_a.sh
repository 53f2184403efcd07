#!/bin/bash
# round 2, session A: the new config-size parity tests + the reworked bench on one B200
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 --round-profile > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2a_bench_c2.json'))
print({k:d[k] for k in ('value','ms_per_step','proof_sha256','parity')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', d['roofline']['frac'], d['roofline']['whole_step'], d['int_roofline'])
for r in d['per_round']: print(r)
t=d['target_c3']
print({k:t[k] for k in t if k!='per_round'})
for r in t['per_round']: print(r)
PY
