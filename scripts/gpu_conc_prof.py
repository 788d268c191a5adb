"""Per-tensor phase times (CUDA events inside the library) with several decompositions in flight."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
from tntorch_b200 import ops

PB = int(sys.argv[1]) if len(sys.argv) > 1 else 4
conc = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
reserve = int(sys.argv[3]) if len(sys.argv) > 3 else (4 if conc else 0)
shape = (64,) * 5
dev = torch.device("cuda:0")
ops.set_reserved_sms(reserve)
Xs = [torch.randn(shape, device=dev) for _ in range(PB)]
plans = [ops.TTSVDPlan(shape, torch.float32, rmax=32, device=dev, profile=True, concurrent=conc) for _ in range(PB)]
streams = [torch.cuda.Stream(device=dev) for _ in range(PB)]
pool = ThreadPoolExecutor(PB)

def run_one(b):
    torch.cuda.set_device(0)
    with torch.cuda.stream(streams[b]):
        return plans[b].run(Xs[b])

def step():
    cur = torch.cuda.current_stream()
    for sb in streams: sb.wait_stream(cur)
    list(pool.map(run_one, range(PB)))
    for sb in streams: cur.wait_stream(sb)

for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"PB={PB} concurrent={conc} reserve={reserve}: {ms:.2f} ms/step  {PB * 64**5 / ms / 1e6:.1f} GElem/s")
for b in range(PB):
    a = np.array(list(plans[b].info))
    n = int(a[7])
    print(f" tensor {b}: gram {[round(a[8+3*s],2) for s in range(n)]} eig {[round(a[9+3*s],2) for s in range(n)]} factor {[round(a[10+3*s],2) for s in range(n)]}  sum {a[4]+a[5]+a[6]:.2f}")
