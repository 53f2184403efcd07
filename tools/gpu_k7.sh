#!/bin/bash
# 2 GPUs: whole -m gpu suite (multi-GPU tests included), smoke under ncu, the default bench line under torchrun
cd "$GRAFT_REPO_ROOT"
bash tools/gpu_validate_2gpu.sh
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/k7_bench_c2_g2.json 2> gpurun_out/k7_bench_c2_g2.err; echo "bench g2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/k7_bench_c2_g2.json') if l.startswith('{')][-1])
print('g2 value %.3e ms %.4f e2e %s parity %s sha %s'%(d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['parity']['ok'], d['proof_sha256'][:16]))
t=d['target_c3']; print(' target ms', t['ms'], t['proof_sha256'][:16], t['oracle_verified']['ok'])
print([round(r['us'],1) for r in d['per_round']])
print([round(r['us'],1) for r in t['per_round']])
PY
