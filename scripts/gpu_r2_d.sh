#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py -x -q 2>&1 | tail -30 | tee gpurun_out/r2d_spec.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_spec.py 2>&1 | tail -15 | tee gpurun_out/r2d_tests.log
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
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2d_launches.csv python /tmp/one.py > gpurun_out/r2d_one.log 2>&1
tail -3 gpurun_out/r2d_one.log
python scripts/launch_summary.py gpurun_out/r2d_launches.csv
python - <<'P'
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
    print('jacobi2 fp64 n',n,'sweeps',sw,'us per call (incl. alloc/launch)',e0.elapsed_time(e1)/20*1000)
P
timeout 600 python bench.py --steps 5 --warmup 3 --per-gpu-batch 1 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench1.json 2> gpurun_out/r2d_bench1.err; echo "bench1 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2d_bench1.json'))
print('single in flight:', d['value'], 'ms', d['ms_per_step'], d['phases_ms'], 'launches', d['gpu_launches'])
P
