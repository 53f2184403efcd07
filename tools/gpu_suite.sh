#!/bin/bash
# round 2, session I: the whole GPU suite + the bench lines for the record (1 GPU)
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/suite_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/suite_pytest.log
grep -v "site-packages" gpurun_out/suite_pytest.log | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --round-profile > gpurun_out/suite_bench_c2.json 2> gpurun_out/suite_bench_c2.err; echo "bench rc=$?"; tail -c 600 gpurun_out/suite_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/suite_bench_ref.json 2>/dev/null; echo "ref rc=$?"
for wl in c1 c3 c4 c4b c5; do
    timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/suite_bench_$wl.json 2>gpurun_out/suite_bench_$wl.err; echo "$wl rc=$?"
done
python - <<'PY'
import json
for wl in ('c2','c1','c3','c4','c4b','c5'):
    try:
        d=json.loads([l for l in open('gpurun_out/suite_bench_%s.json'%wl) if l.startswith('{')][-1])
        e=d.get('e2e') or {}
        print(wl, 'value %.3e ms/step %.4f e2e %.3e launches %s roofline %s' % (d['value'], d['ms_per_step'], e.get('value',0), d['gpu_launches'], (d.get('roofline') or {}).get('frac')))
        if wl=='c2':
            print(' parity', d['parity']['ok'], 'sha', d['proof_sha256'][:16], 'whole', d['roofline']['whole_step'], 'int', d['int_roofline']['peak'])
            t=d['target_c3']; print(' target', t['ms'], t['proof_sha256'][:16], t['oracle_verified']['ok'], t['whole_step'], t['rounds_at_or_above_0.60'])
            print(' cpu', d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:90])
    except Exception as ex:
        print(wl, 'FAILED', ex)
r=json.load(open('gpurun_out/suite_bench_ref.json')); print('ref', r['value'], r['config']['n_vars'], r['cpu_sample_n_vars'], r['ms_per_step'])
PY
