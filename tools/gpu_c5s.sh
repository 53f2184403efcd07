#!/bin/bash
# BASELINE config 5's gather leg: the 64 x 2^22 batch with every proof sharded over N GPUs (bench.py --workload c5s), next to the
# replicas form (c5); usage: bash tools/gpu_c5s.sh <N>
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
for wl in c5s c5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2981$N bench.py --gpus $N --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${wl}_g$N.json 2> gpurun_out/${wl}_g$N.err; echo "$wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${wl}_g$N.json') if l.startswith('{')][-1])
    print('$wl N=$N value %.3e ms/step %.3f launches %d sha %s sharding: %s' % (d['value'], d['ms_per_step'], d['gpu_launches'], d['proof_sha256'][:12], d['config']['sharding'][:40]))
except Exception as e:
    print('$wl FAILED', e); print(open('gpurun_out/${wl}_g$N.err').read()[-1500:])
PY
done
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('c5 N=1 value %.3e ms/step %.3f sha %s' % (d['value'], d['ms_per_step'], d['proof_sha256'][:12]))"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final2_pytest.log 2>&1; tail -3 gpurun_out/final2_pytest.log
