#!/bin/bash
# device-side timelines of the resident kernel (debug build in build/trace) for degree 2 and 3
cd "$GRAFT_REPO_ROOT"
export ZKSC_LIB=$PWD/build/trace/libzksc_trace.so ZKSC_TRACE_CTAS=1
for a in "24 2" "24 3"; do
  echo "== n, degree = $a"; python tools/trace_resident.py $a 2>&1 | tail -60
done > gpurun_out/trace_resident_k.txt
cat gpurun_out/trace_resident_k.txt | grep -v "^ *[0-9]* |"
