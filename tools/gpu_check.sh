#!/bin/bash
# Quick GPU check: parity tests, smoke, benches.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
for w in ${WORKLOADS:-c2}; do
python bench.py --workload $w --steps 5 --warmup 3 ${BENCH_FLAGS} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 3500 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
