#!/bin/bash
# round 2, pass M: project_tc with V resident up to 128 KB and early stage hand-back (CP-ALS projections), A/B by env;
# cfg4 per-sweep time with a warm workspace; TT bench sanity with the 64-row gram_tc2 stages
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_cp.py tests/test_gpu_callers.py -m gpu -x -q 2>&1 | tail -3
cat > /tmp/pj.py <<'P'
import sys, torch
sys.path.insert(0, ".")
from tntorch_b200 import ops
def t(rows, K, r, n=5):
    A = torch.randn(rows, K, device="cuda"); V = torch.randn(K, r, device="cuda")
    for _ in range(2): ops.project(A, V, tensorcore=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): ops.project(A, V, tensorcore=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"project {rows}x{K} r={r}: {ms:.3f} ms, {(rows*K*4 + rows*r*4)/ms/1e6:.0f} GB/s", flush=True)
t(256**3, 256, 50); t(64**4, 64, 32); t(64**3, 2048, 32); t(256**3 // 2, 512, 32)
P
for cfg in "" "TNB_PT_VRES_KB=32" "TNB_PT_EARLY=0" "TNB_PT_EARLY=1"; do
  echo "== env: $cfg"
  env $cfg python /tmp/pj.py
done
python scripts/bench_extra.py cfg4 2>&1 | tail -1 | tee gpurun_out/r02_cfg4.json
TNB_PT_VRES_KB=32 python scripts/bench_extra.py cfg4 2>&1 | tail -1
python scripts/batch_exp.py 8 4
