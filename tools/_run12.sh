set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg4.csv python tools/run_configs.py 4 --repeats 1 > gpurun_out/r2_ncu_cfg4.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2_launches_cfg4.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ix={k:j for j,k in enumerate(h)}
agg=collections.OrderedDict()
for r in rows[start+2:]:
    try: name=r[ix['Kernel Name']].split('(')[0][:70]; t=float(r[ix['Metric Value']].replace(',',''))
    except: continue
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"  {t/1e6:9.3f} ms  x{c:4d}  {k}")
PY
