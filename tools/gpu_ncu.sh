#!/bin/bash
# ncu full captures (source-level) of the round kernels on c2; one launch each.
mkdir -p gpurun_out
export ZKSC_NO_STAGED=${ZKSC_NO_STAGED:-0}
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_.*kernelILi2ELb0 -s 3 -c 1 -f -o gpurun_out/prof_d2_eval_$1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_.*kernelILi2ELb1 -s 3 -c 1 -f -o gpurun_out/prof_d2_fold_$1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu >> gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
