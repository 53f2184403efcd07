#!/bin/bash
# Multi-GPU session: parity against one GPU and the oracle, then the sharded benches.  Usage: gpu_multi.sh G
G=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$G.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/multi_check_$G.txt 2>&1
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/multi_check_$G.txt | head -60
for w in ${WORKLOADS:-c2 c3}; do
timeout 600 $TR bench.py --gpus $G --workload $w --steps 5 --warmup 3 --round-profile ${BENCH_FLAGS} > gpurun_out/bench_${w}_g$G.json 2> gpurun_out/bench_${w}_g$G.err
tail -c 2500 gpurun_out/bench_${w}_g$G.json; tail -3 gpurun_out/bench_${w}_g$G.err
done
