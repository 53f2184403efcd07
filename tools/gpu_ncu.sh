#!/bin/bash
# round 2: Nsight Compute evidence (ordinary launches only: under ncu every launch is synchronous, so a resident kernel can never be
# answered by the host and times out -- ZKSC_NO_TAIL=1 keeps the run on the per-round kernels, which are what is being profiled)
cd "$GRAFT_REPO_ROOT"
export ZKSC_NO_TAIL=1
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-target"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c2.csv $B --workload c2 > gpurun_out/r02_launches_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"round_(tma_)?kernel" -s 72 -c 2 -f -o gpurun_out/r02_prof_c2 $B --workload c2 > gpurun_out/r02_prof_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"round_(tma_)?kernel" -s 84 -c 2 -f -o gpurun_out/r02_prof_c3 $B --workload c3 > gpurun_out/r02_prof_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r02_prof_c2.log gpurun_out/r02_prof_c3.log
