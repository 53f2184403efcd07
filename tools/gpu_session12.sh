#!/bin/bash
# host worker pool for batched proofs: parity, c5 with and without the pool, c2 unchanged
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_s12.txt
for tag in pool nopool; do
  if [ $tag = nopool ]; then export ZKSC_HOST_THREADS=1; fi
  python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_c5_$tag.json 2> gpurun_out/bench_c5_$tag.err; tail -2 gpurun_out/bench_c5_$tag.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c5_$tag.json").read().strip().splitlines()[-1])
print("$tag c5 value %.3f G, %.3f ms/step, e2e %.3f" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9))
PY
done
unset ZKSC_HOST_THREADS
python bench.py --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value']/1e9, d['ms_per_step'])"
