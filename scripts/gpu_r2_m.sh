#!/bin/bash
# round 2, pass M2: project_tc epilogue for r % 4 != 0 staged through shared memory (coalesced stores); which part of the
# K = 256, r = 50 projection is slow (r = 48 / 50 / 52 / 64); cfg4 per-sweep time with a warm workspace
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_cp.py tests/test_gpu_callers.py -m gpu -q 2>&1 | tail -5
cat > /tmp/pj.py <<'P'
import sys, torch
sys.path.insert(0, ".")
from tntorch_b200 import ops
def t(rows, K, r, n=5):
    A = torch.randn(rows, K, device="cuda"); V = torch.randn(K, r, device="cuda")
    for _ in range(2): C = ops.project(A, V, tensorcore=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): ops.project(A, V, tensorcore=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    ref = A[:4096].double() @ V.double()
    err = float((C[:4096].double() - ref).abs().max() / ref.abs().max())
    print(f"project {rows}x{K} r={r}: {ms:.3f} ms, {(rows*K*4 + rows*r*4)/ms/1e6:.0f} GB/s, err {err:.1e}", flush=True)
for r in (32, 48, 50, 52, 64): t(256**3, 256, r)
t(64**4, 64, 32); t(64**3, 2048, 32); t(64**4, 64, 30)
P
python /tmp/pj.py
python scripts/bench_extra.py cfg4 2>&1 | tail -1 | tee gpurun_out/r02_cfg4.json
python scripts/batch_exp.py 8 4
