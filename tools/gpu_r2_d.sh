#!/bin/bash
cd "$GRAFT_REPO_ROOT"
ZKSC_LIB=$PWD/build/libzksc_trace.so timeout 200 python tools/trace_resident.py 24 2 > gpurun_out/r2d_trace_c2.txt 2>&1; cat gpurun_out/r2d_trace_c2.txt
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/r2d_c2.json 2>gpurun_out/r2d_c2.err
python -c "
import json
d=json.load(open('gpurun_out/r2d_c2.json')); print('c2 ms/step %.4f'%d['ms_per_step'], [round(r['us'],1) for r in d['per_round']])"
