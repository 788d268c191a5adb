"""Measure the dense TF32 tcgen05 peak of this GPU and write profiles/r02_tf32_peak.json (read by bench.py)."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tntorch_b200 import ops

torch.cuda.set_device(0)
best = 0.0
rows = []
for per_commit in (16, 64, 256):
    tf, ms = ops.measure_tf32_peak(reps=2048 // per_commit * 8, per_commit=per_commit, trials=5)
    rows.append({"per_commit": per_commit, "tflops": tf, "ms": ms})
    best = max(best, tf)
out = {"tf32_tflops": best, "runs": rows, "gpu": torch.cuda.get_device_name(0),
       "how": "csrc/peak_tf32.cuh: 148 CTAs x tcgen05.mma.cta_group::1.kind::tf32 M=128 N=256 K=8 on smem-resident tiles, best of 5, CUDA events"}
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "r02_tf32_peak.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
