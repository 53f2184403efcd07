#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  for w in ${WORKLOADS:-c2}; do
    env "$@" python bench.py --workload $w --steps 8 --warmup 3 --no-cpu --no-e2e --round-profile > gpurun_out/var_${name}_$w.json 2> gpurun_out/var_${name}_$w.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${name}_$w.json"))
    rp=d.get("round_profile",[])[:5]
    print("${name} $w: %.3f G evals/s, %.3f ms/step | "%(d["value"]/1e9,d["ms_per_step"]), [(r["fold"],r["pairs"],r["avg_us"]) for r in rp])
except Exception as e:
    print("${name} $w: FAILED", e); print(open("gpurun_out/var_${name}_$w.err").read()[-800:])
PY
  done
}
run plain ZKSC_X=0
for n in ${VARIANTS}; do run $n ZKSC_LIB=$PWD/build/libzksc_$n.so; done
run plain_again ZKSC_X=0
