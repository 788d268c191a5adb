#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/t_all.log | cut -c1-250
python - <<'PY'
import torch, time, sys, os
sys.path.insert(0, os.getcwd())
import tntorch_b200 as tn
from oracle import cases
for name in ("eps_biggram_f32", "eps_biggram_f64"):
    X = torch.as_tensor(cases.make_dense(cases.TTSVD_CASES[name])).cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t = tn.Tensor(X, eps=cases.TTSVD_CASES[name]["eps"])
    torch.cuda.synchronize(); print(name, list(t.ranks_tt), [None if u is None else tuple(u.shape) for u in t.Us], f"{(time.perf_counter()-t0)*1e3:.1f} ms")
PY
