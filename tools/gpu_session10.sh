#!/bin/bash
# Final measurement pass of this session (one B200): parity, smoke, reference arm + default bench, the other workloads,
# ncu launch list of the default bench command and --set full captures of the two dominant c2 kernels as they are now.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_s10.txt
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
python bench.py --round-profile > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
for w in c1 c3 c4 c5; do python bench.py --workload $w --steps 5 --warmup 3 --round-profile > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; done
python - <<'PY'
import json
for w in ["c2", "c1", "c3", "c4", "c5"]:
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % w).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(w, "value %.3f G, %.3f ms/step, e2e %s, roofline %s binding %s" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"] / 1e9, 3), r.get("frac"), (r.get("binding") or {}).get("frac")))
    except Exception as e:
        print(w, "FAILED", e)
PY
export ZKSC_NO_TAIL=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
export ZKSC_NO_MAPPED=1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_kernelILi2ELb1ELb1ELi1 -s 0 -c 1 -f -o gpurun_out/prof_d2_fold_v6 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:round_tma_kernelILi2ELb0 -s 0 -c 1 -f -o gpurun_out/prof_d2_eval_v6 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu >> gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*_v6.ncu-rep
