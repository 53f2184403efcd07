#!/bin/bash
# after the ALU trimming of the fold tail / last evaluation point: parity first, then the per-round profile of c2, c3, c5
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_s9.txt
WORKLOADS="c2 c3 c5" VARIANTS="" bash tools/gpu_variants3.sh 2>&1 | grep -v plain_again
