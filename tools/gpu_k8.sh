#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_gkr.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
ZKSC_AB_WORKLOADS="c2 c3 c1 c5" bash tools/gpu_ab_env.sh "k8:ZKSC_X=0"
export ZKSC_LIB=$PWD/build/trace/libzksc_trace.so ZKSC_TRACE_CTAS=1
for a in "24 2" "24 3"; do
  echo "== n, degree = $a"; python tools/trace_resident.py $a 2>&1 | tail -60
done > gpurun_out/trace_resident_k8.txt
grep -v "earliest\|latest\|SMs by\|duration\|decile" gpurun_out/trace_resident_k8.txt
