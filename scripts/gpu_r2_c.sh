#!/bin/bash
# round 2, call C: launch list of one speculative 64^5 decomposition + Jacobi micro-benchmarks
mkdir -p gpurun_out
cat > /tmp/one.py <<'P'
import torch, sys
sys.path.insert(0, '.')
from tntorch_b200 import ops
g = torch.Generator(device="cuda").manual_seed(7)
X = torch.randn((64,) * 5, generator=g, device="cuda")
plan = ops.TTSVDPlan((64,)*5, torch.float32, rmax=32)
for _ in range(2): plan.run(X)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
plan.run(X)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(list(plan.info)[:8], plan.info[26], plan.info[27], plan.info[2], plan.info[30])
P
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches.csv python /tmp/one.py > gpurun_out/r2c_one.log 2>&1
tail -3 gpurun_out/r2c_one.log
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2c_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[]
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    us = v/1000 if u in('ns','nsecond') else (v if u in ('us','usecond') else v*1000)
    seq.append((r[ki][:60],us))
agg=collections.OrderedDict()
for k,u in seq:
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=u
print('launches',len(seq),'total us',sum(u for _,u in seq))
for k,(c,u) in sorted(agg.items(), key=lambda x:-x[1][1])[:25]: print(f'{u:10.1f} us  x{c:4d}  {k}')
P
python - <<'P'
import torch, time, os, sys
sys.path.insert(0,'.')
from tntorch_b200 import ops
import numpy as np
rng=np.random.default_rng(0)
for n in (32,64,80):
    A=rng.standard_normal((n+3,n)); G=torch.as_tensor(A.T@A).cuda()
    for rep in range(2):
        w,V,sw=ops.eigh_jacobi(G,return_sweeps=True)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.eigh_jacobi(G)
    e1.record(); torch.cuda.synchronize()
    print('jacobi2 fp64 n',n,'sweeps',sw,'us per call (incl. alloc/launch)',e0.elapsed_time(e1)/20*1000)
P
TNB_NO_JACOBI2=1 python - <<'P'
import torch, sys
sys.path.insert(0,'.')
from tntorch_b200 import ops
import numpy as np
rng=np.random.default_rng(0)
for n in (32,64,80):
    A=rng.standard_normal((n+3,n)); G=torch.as_tensor(A.T@A).cuda()
    for rep in range(2):
        w,V,sw=ops.eigh_jacobi(G,return_sweeps=True)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.eigh_jacobi(G)
    e1.record(); torch.cuda.synchronize()
    print('jacobi (old) fp64 n',n,'sweeps',sw,'us per call',e0.elapsed_time(e1)/20*1000)
P
