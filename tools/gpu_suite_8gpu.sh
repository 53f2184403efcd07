#!/bin/bash
# round 2 (8 GPUs): multi-process sharded parity at 8 and 4 ranks, the single-process multi-GPU context at 2/4/8 devices, and the
# bench lines at N = 8 and 4 (c2 weak-scaling headline + the target_c3 leg: 2^28, degree 3, strong scaling, oracle-verified)
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L | wc -l
ZKSC_CHECK_MAX_N=26 timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/suite8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/suite8_pytest.log
grep -v "^E  \|site-packages" gpurun_out/suite8_pytest.log | grep "case n=2[2-8]\|PARITY\|passed\|failed\|rc=" | tail -20
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/suite8_bench_g$N.json 2> gpurun_out/suite8_bench_g$N.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/suite8_bench_g$N.json') if l.startswith('{')][-1])
print('N=$N c2 weak: value %.3e ms %.4f e2e %.3e sha %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['proof_sha256'][:12], d['parity']['ok']))
print([ (r['pairs'], round(r['us'],1)) for r in d['per_round']])
t=d['target_c3']
print('target:', {k:t[k] for k in t if k not in ('per_round','workload','per_round_clock','oracle_verified')}, t['oracle_verified']['ok'])
print([ (r['pairs'], round(r['us'],1), r['frac']) for r in t['per_round']])
PY
done
