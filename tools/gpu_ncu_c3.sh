#!/bin/bash
# ncu --set full captures of the two dominant c3 kernels (d = 3 fold with shared-memory sums, staged round 0), one launch each
mkdir -p gpurun_out
export ZKSC_NO_TAIL=1 ZKSC_NO_MAPPED=1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_kernelILi3ELb1ELb1ELi1 -s 0 -c 1 -f -o gpurun_out/prof_d3_fold_v8 python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_tma_kernelILi3ELb0 -s 0 -c 1 -f -o gpurun_out/prof_d3_eval_v8 python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu >> gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log; ls -la gpurun_out/*d3*_v8.ncu-rep
