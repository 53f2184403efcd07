#!/bin/bash
# GKR C driver: parity tests, timing (C call vs the Python layer driver), c4 bench line.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gkr.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gkr.txt
python tools/profile_gkr.py 2>&1 | head -12 | tee gpurun_out/profile_gkr.txt
ZKSC_PROFILE=1 python - <<'PY' 2>&1 | tail -25 | tee gpurun_out/gkr_rounds.txt
import sys; sys.path.insert(0, '.')
import zk_cryptography_b200 as zk
from bench import gkr_inputs
ctx = zk.Context(0); zk.set_default_context(ctx)
c = zk.Circuit.random(10); ev = c.evaluation(gkr_inputs(10))
inst = zk.GKRInstance(c, ev)
for _ in range(2): inst.prove_raw(ctx)
PY
python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 2500 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
