#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/bench_extra.py > gpurun_out/extra.log 2>&1; tail -n 1 gpurun_out/extra.log | cut -c1-1500
cat > /tmp/proj2.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from tntorch_b200 import ops
A = torch.randn(1 << 17, 2048, device="cuda"); V = torch.randn(2048, 32, device="cuda")
for _ in range(3):
    C = ops.project(A, V, tensorcore=True)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_tc_kernel -c 1 -s 2 -o gpurun_out/prof_proj2 python /tmp/proj2.py > gpurun_out/ncu_proj2.log 2>&1; echo "ncu rc=$?"
