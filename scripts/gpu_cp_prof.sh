#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/cp_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from tntorch_b200 import ops
shape = (128,) * 4
g = torch.Generator(device="cuda").manual_seed(1)
fs = [torch.randn(s, 50, generator=g, device="cuda") for s in shape]
X = torch.einsum("ar,br,cr,dr->abcd", *fs)
fac, info = ops.cp_als(X, 50, max_iter=2, tol=float("-inf"), return_info=True)
torch.cuda.synchronize(); print(info)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cp_launches.csv python /tmp/cp_one.py > gpurun_out/cp_ncu.log 2>&1; echo rc=$?
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/cp_launches.csv') if l.startswith('"')]
agg=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(lines):
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    name=re.sub(r'\(.*','',r['Kernel Name'])[:90]
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    v = v/1e6 if u=='ns' else v/1e3 if u in('us','usecond') else v
    agg[name][0]+=1; agg[name][1]+=v
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print(f"{v[1]:9.3f} ms {v[0]:5d}  avg {v[1]/v[0]*1e3:9.1f} us  {k}")
PY
