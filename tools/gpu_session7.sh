#!/bin/bash
# GKR C driver + double-buffered refill: parity tests, per-layer profile, c4 / c2 / c5 bench lines.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gkr.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_s7.txt
ZKSC_PROFILE=1 python - <<'PY' 2>&1 | grep "gkr layer" | tail -10 | tee gpurun_out/gkr_layers.txt
import sys; sys.path.insert(0, '.')
import zk_cryptography_b200 as zk
from bench import gkr_inputs
ctx = zk.Context(0); zk.set_default_context(ctx)
c = zk.Circuit.random(10); ev = c.evaluation(gkr_inputs(10))
inst = zk.GKRInstance(c, ev)
for _ in range(3): inst.prove_raw(ctx)
PY
python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 600 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python - <<'PY'
import json; d = json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], json.dumps(d["e2e"]))
PY
tail -3 gpurun_out/bench_c2.err
python bench.py --workload c5 --steps 5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python - <<'PY'
import json; d = json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], json.dumps(d["e2e"]), json.dumps(d["roofline"]["binding"]))
PY
tail -3 gpurun_out/bench_c5.err
