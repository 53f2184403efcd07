#!/bin/bash
# robustness of the resident-kernel paths: smoke under Nsight Compute (synchronous launches), then the stress run on one GPU
cd "$GRAFT_REPO_ROOT"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/robust_smoke_ncu.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/robust_smoke_ncu.log 2>&1; echo "ncu smoke rc=$?"; tail -2 gpurun_out/robust_smoke_ncu.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/robust_smoke_ncu.csv')) if len(r)>5 and r[0].isdigit()]
c=collections.Counter(); t=collections.Counter()
for r in rows:
    name=r[4][:60]; c[name]+=1
    try: t[name]+=float(r[-1].replace(',',''))
    except: pass
for k,v in c.most_common(): print(v, k, t[k])
PY
timeout 300 python tools/stress.py --seconds 50 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
