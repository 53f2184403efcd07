#!/bin/bash
# ncu session (one GPU): launch list of the default bench command, then --set full captures of the two dominant
# round kernels of c2 (one launch each).  The full captures run with ZKSC_NO_TAIL=1 ZKSC_NO_MAPPED=1: kernel replay
# cannot re-run a resident kernel that talks to the host, and the round kernels are identical either way.  The launch
# list also runs with ZKSC_NO_TAIL=1: under ncu a launch returns only when the kernel has ended, so a resident kernel
# waiting for the host can only time out (the library then falls back to one launch per round by itself).
mkdir -p gpurun_out
export ZKSC_NO_TAIL=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
export ZKSC_NO_TAIL=1 ZKSC_NO_MAPPED=1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_kernelILi2ELb1ELb1ELi1 -s 0 -c 1 -f -o gpurun_out/prof_d2_fold python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_tma_kernelILi2ELb0 -s 0 -c 1 -f -o gpurun_out/prof_d2_eval python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu >> gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
