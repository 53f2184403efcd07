#!/bin/bash
# 2 GPUs: the multi-GPU tests and the default bench line under torchrun (what the driver's scaling step runs)
cd "$GRAFT_REPO_ROOT"
ZKSC_CHECK_MAX_N=24 timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final_bench_c2_g2.json 2> gpurun_out/final_bench_c2_g2.err; echo "bench g2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final_bench_c2_g2.json') if l.startswith('{')][-1])
print('g2 value %.3e ms %.4f e2e %s parity %s sha %s'%(d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['parity']['ok'], d['proof_sha256'][:16]))
t=d['target_c3']; print(' target ms', t['ms'], t['proof_sha256'][:16], t['oracle_verified']['ok'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference 2>/dev/null | tail -1 | cut -c1-200
