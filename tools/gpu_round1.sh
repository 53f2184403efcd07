#!/bin/bash
# Round-1 GPU session: parity tests, bench, ncu launch list + full capture of the top kernel, microbenchmarks.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi0.txt
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_kernelILi2ELb1 -s 0 -c 2 -o gpurun_out/prof_d2_fold python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_kernelILi2ELb0 -s 0 -c 1 -o gpurun_out/prof_d2_eval python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu >> gpurun_out/ncu_full.log 2>&1
build/ubench2 > gpurun_out/ubench2.jsonl; cat gpurun_out/ubench2.jsonl
