#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches.csv') if l.startswith('"')]
rd=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0])
for r in rd:
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    name=re.sub(r'\(.*','',r['Kernel Name'])[:90]
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    v = v/1e6 if u=='ns' else v/1e3 if u in('us','usecond') else v
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values()); print("total ms",tot)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]:
    print(f"{v[1]:9.3f} ms {v[0]:6d} launches  avg {v[1]/v[0]*1e3:9.1f} us  {k}")
PY
