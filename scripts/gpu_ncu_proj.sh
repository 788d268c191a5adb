#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/proj_only.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from tntorch_b200 import ops
for rows, K in ((1 << 22, 64), (1 << 17, 2048)):
    A = torch.randn(rows, K, device="cuda"); V = torch.randn(K, 32, device="cuda")
    for _ in range(2):
        C = ops.project(A, V, tensorcore=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        C = ops.project(A, V, tensorcore=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("rows", rows, "K", K, "ms", ms, "GB/s", (A.numel() + C.numel()) * 4 / ms / 1e6)
PY
timeout 300 python /tmp/proj_only.py
timeout 900 ncu --set full --clock-control none --import-source on -k regex:project_tc_kernel -c 2 -s 1 -o gpurun_out/prof_proj python /tmp/proj_only.py > gpurun_out/ncu_proj.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/ncu_proj.log
ls -la gpurun_out/prof_proj.ncu-rep
