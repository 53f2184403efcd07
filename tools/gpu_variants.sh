#!/bin/bash
# Kernel A/B on one box: same bench, alternative builds of libzksc.so / env switches.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
run() { # name, env...
  name=$1; shift
  for w in ${WORKLOADS:-c2 c3}; do
    env "$@" python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e --round-profile > gpurun_out/var_${name}_$w.json 2> gpurun_out/var_${name}_$w.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${name}_$w.json"))
    rp=d.get("round_profile",[])[:4]
    print("${name} $w: %.3f G evals/s, %.3f ms/step, top %s %.1f us frac %.3f | "%(d["value"]/1e9,d["ms_per_step"],d["roofline"]["kernel"],d["roofline"]["launch"]["avg_ms"]*1e3,d["roofline"]["frac"]), [(r["fold"],r["pairs"],r["avg_us"]) for r in rp])
except Exception as e:
    print("${name} $w: FAILED", e); print(open("gpurun_out/var_${name}_$w.err").read()[-800:])
PY
  done
}
run staged_default ZKSC_X=1
run staged_minb3 ZKSC_LIB=$PWD/build/libzksc_minb3.so
run unstaged ZKSC_NO_STAGED=1
