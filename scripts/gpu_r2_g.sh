#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cp.py tests/test_gpu_spec.py tests/test_gpu_batch.py tests/test_gpu_fullgolden.py tests/test_gpu_blocks.py -x -q -s 2>&1 | tail -25 | tee gpurun_out/r2g_tests.log
for order in stage phase; do
  echo "== order=$order"
  TNB_BATCH_ORDER=$order timeout 300 python scripts/batch_exp.py 6 4 2>&1 | tail -1
done
TNB_NO_GATE=1 timeout 300 python scripts/batch_exp.py 6 4 2>&1 | tail -1
timeout 300 python scripts/batch_exp.py 4 4 2>&1 | tail -1
timeout 300 python scripts/batch_exp.py 8 4 2>&1 | tail -1
timeout 300 python scripts/batch_exp.py 6 8 2>&1 | tail -1
timeout 300 python scripts/batch_exp.py 1 0 2>&1 | tail -1
python - <<'P'
import torch, sys
sys.path.insert(0,'.')
from tntorch_b200 import ops
import numpy as np
rng=np.random.default_rng(0)
for n in (32,64):
    A=rng.standard_normal((n+3,n)); G=torch.as_tensor(A.T@A).cuda()
    for rep in range(2):
        w,V,sw=ops.eigh_jacobi(G,return_sweeps=True)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.eigh_jacobi(G)
    e1.record(); torch.cuda.synchronize()
    print('jacobi2 mixed fp64 n',n,'sweeps(32+64)',sw,'us per call (incl. alloc/launch)',e0.elapsed_time(e1)/20*1000)
P
python scripts/bench_extra.py cfg3 cfg5 2>&1 | tail -1
