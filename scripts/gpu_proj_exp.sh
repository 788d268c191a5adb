#!/bin/bash
cat > /tmp/proj_tr.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from tntorch_b200 import ops
rows, K = (1 << 22, 64)
A = torch.randn(rows, K, device="cuda"); V = torch.randn(K, 32, device="cuda")
for _ in range(4):
    C = ops.project(A, V, tensorcore=True)
torch.cuda.synchronize()
PY
TNB_PT_TRACE=1 timeout 120 python /tmp/proj_tr.py 2>&1 | tail -n 70
