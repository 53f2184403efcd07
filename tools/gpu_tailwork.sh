#!/bin/bash
# start threshold of the resident rounds kernel: limb products per round per CTA of a group (default 460000)
mkdir -p gpurun_out
for tw in 460000 920000 1840000 3680000; do
  for w in c2 c1 c4 c5; do
    ZKSC_TAIL_WORK=$tw python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/tw.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tail_work $tw $w: %.3f G evals/s, %.4f ms/step, launches %d' % (d['value']/1e9, d['ms_per_step'], d['gpu_launches']))"
  done
done
