#!/bin/bash
# round 2, pass P: float2 Khatri-Rao reduction — CP tests, cfg4 per-sweep time, per-launch times
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cp.py tests/test_gpu_fullgolden.py -m gpu -q 2>&1 | tail -3
python scripts/bench_extra.py cfg4 2>&1 | tail -1 | tee gpurun_out/r02_cfg4.json
cat > /tmp/cp2.py <<'P'
import sys, torch
sys.path.insert(0, ".")
from tntorch_b200 import ops
X = torch.randn(256, 256, 256, 256, device="cuda")
ops.cp_als(X, 50, max_iter=1, tol=float("-inf"))
torch.cuda.synchronize()
P
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:khatri -c 7 --csv --log-file gpurun_out/r02_khatri_after.csv python /tmp/cp2.py > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_khatri_after.csv | head -4
