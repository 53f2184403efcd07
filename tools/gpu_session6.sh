#!/bin/bash
# Session-6 GPU check (one B200): parity tests, smoke, default bench + reference arm, then the ncu launch list of the
# default bench command and --set full captures of the dominant kernels as they are now (128-register cap).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 4000 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
export ZKSC_NO_TAIL=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
export ZKSC_NO_MAPPED=1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_kernelILi2ELb1ELb1ELi1 -s 0 -c 1 -f -o gpurun_out/prof_d2_fold_v5 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_kernelILi3ELb1ELb1ELi1 -s 0 -c 1 -f -o gpurun_out/prof_d3_fold_v5 python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu >> gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
