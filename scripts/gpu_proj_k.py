"""Projection kernel bandwidth vs K (row pitch) at fixed bytes, r = 32: isolates what limits the K = 2048 case."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tntorch_b200 import ops
total = 1 << 28  # elements of A (1 GiB)
for K in (64, 128, 256, 512, 1024, 2048, 4096):
    rows = total // K
    A = torch.randn(rows, K, device="cuda"); V = torch.randn(K, 32, device="cuda")
    for _ in range(3): C = ops.project(A, V, tensorcore=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): C = ops.project(A, V, tensorcore=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"K={K:5d} rows={rows:8d} V {'resident' if K <= 128 else 'streamed'}: {ms:.3f} ms  {(A.numel() + C.numel()) * 4 / ms / 1e6:7.0f} GB/s", flush=True)
    del A, V, C
