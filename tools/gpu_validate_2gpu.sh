cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/v2_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v2_pytest_2gpu.log
grep -v "site-packages\|^E  " gpurun_out/v2_pytest_2gpu.log | tail -16
# the driver traces smoke() under Nsight Compute: every launch is synchronous there, the resident kernel must give up cleanly and the proof must still be right
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v2_smoke_ncu.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v2_smoke_ncu.log 2>&1; echo "ncu smoke rc=$?"; tail -2 gpurun_out/v2_smoke_ncu.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/v2_smoke_ncu.csv')) if len(r)>5 and r[0].isdigit()]
c=collections.Counter(); t=collections.Counter()
for r in rows:
    name=r[4][:60]; c[name]+=1
    try: t[name]+=float(r[-1].replace(',',''))
    except: pass
for k,v in c.most_common(): print(v, k, t[k])
PY
