"""Times the resident Chebyshev filter (one kernel, `steps` products) against product-at-a-time launches."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tntorch_b200 import ops

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for n, b in [(512, 64), (1024, 64), (1536, 64), (2048, 64), (2048, 32)]:
    A = torch.randn(n, n, device="cuda")
    G = (A @ A.T / n).contiguous()
    Y = torch.randn(n, b, device="cuda")
    for steps in (1, 20, 40):
        a = [0.5 / n ** 0.5] * steps; bc = [-0.3] * steps; g = [0.0] + [-0.25] * (steps - 1)
        for mode in ("dsmem", "l2"):
            if mode == "l2": os.environ["TNB_FILTER_NO_DSMEM"] = "1"
            else: os.environ.pop("TNB_FILTER_NO_DSMEM", None)
            try:
                ms = timeit(lambda: ops.cheb_filter(G, Y, a, bc, g))
            except Exception as e:
                print(n, b, steps, mode, "ERR", str(e)[:100]); continue
            print(f"n={n} b={b} steps={steps} {mode}: {ms*1e3:.1f} us total, {ms*1e3/steps:.2f} us/step", flush=True)
    ms = timeit(lambda: ops.atb_tensorcore(G, Y, 1.0, Y, 0.5))
    print(f"n={n} b={b} single product launch pair: {ms*1e3:.1f} us")
