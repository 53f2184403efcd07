#!/bin/bash
# host-side round step (SHA-NI, cached Lagrange basis / constants): parity, per-round host profile, bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_s11.txt
ZKSC_PROFILE=1 python tools/profile_rounds.py 2>&1 | tail -26 | tee gpurun_out/profile_rounds_c2.txt
for w in c2 c1 c4 c5; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; done
python - <<'PY'
import json
for w in ["c2", "c1", "c4", "c5"]:
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % w).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(w, "value %.3f G, %.3f ms/step, e2e %s, roofline %s" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"] / 1e9, 3), r.get("frac")))
    except Exception as e:
        print(w, "FAILED", e)
PY
