#!/bin/bash
# round 2, session E (2 GPUs): sharded parity incl. the in-kernel gather, then the bench at N=2
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -40 gpurun_out/r2e_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_g2.json 2> gpurun_out/r2e_bench_g2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2e_bench_g2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2e_bench_g2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','proof_sha256','parity','exchange')})
print([ (r['pairs'], round(r['us'],1)) for r in d['per_round']])
t=d['target_c3']
print({k:t[k] for k in t if k!='per_round'})
print([ (r['pairs'], round(r['us'],1), r['frac']) for r in t['per_round']])
PY
