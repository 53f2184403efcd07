#!/bin/bash
# compute-sanitizer over a small end-to-end run (tools/sanitize.py); ordinary launches only for racecheck (a resident kernel spins on host memory)
cd "$GRAFT_REPO_ROOT"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds" gpurun_out/sanitize_memcheck.log; tail -6 gpurun_out/sanitize_memcheck.log
ZKSC_NO_TAIL=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/sanitize_racecheck.log
