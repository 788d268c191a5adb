#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/jac_only.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from tntorch_b200 import ops
A = torch.randn(256, 64, dtype=torch.float64, device="cuda"); G = A.T @ A
for _ in range(2):
    w, V = ops.eigh_jacobi(G)
torch.cuda.synchronize()
print("ok", w[:3].tolist())
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_eigh_kernel -c 1 -s 1 -o gpurun_out/prof_jacobi python /tmp/jac_only.py > gpurun_out/ncu_jac.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/ncu_jac.log
