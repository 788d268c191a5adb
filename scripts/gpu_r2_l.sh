#!/bin/bash
# round 2, pass L: the new callers + CP-ALS dimension tree (tests, cfg4 per-sweep time, launch list), cfg3 launch list,
# and the A/B of gram_tc2's pipeline depth (TNB_TC2_KC=32: 4 KB boxes x 6 stages; 64: 8 KB boxes x 3 stages)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_callers.py tests/test_gpu_cp.py tests/test_gpu_fullgolden.py -m gpu -q 2>&1 | tail -8
python scripts/bench_extra.py cfg4 2>&1 | tail -1 | tee gpurun_out/r02_cfg4.json
cat > /tmp/cp1.py <<'P'
import sys, torch
sys.path.insert(0, ".")
from tntorch_b200 import ops
X = torch.randn(256, 256, 256, 256, device="cuda")
ops.cp_als(X, 50, max_iter=2, tol=float("-inf"))
torch.cuda.synchronize()
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_cfg4_launches.csv python /tmp/cp1.py > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_cfg4_launches.csv 14 | tee gpurun_out/r02_cfg4_launch_summary.txt
cat > /tmp/rt1.py <<'P'
import sys, torch
sys.path.insert(0, ".")
import tntorch_b200 as tnb
g = torch.Generator(device="cuda").manual_seed(0)
rs = [1] + [64] * 9 + [1]
cores = [torch.randn(rs[k], 128, rs[k + 1], generator=g, device="cuda", dtype=torch.float64) for k in range(10)]
t = tnb.Tensor(cores)
tnb.round_tt(t, rmax=16)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tnb.round_tt(t, rmax=16)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
P
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/r02_cfg3_launches.csv python /tmp/rt1.py > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_cfg3_launches.csv 14 | tee gpurun_out/r02_cfg3_launch_summary.txt
for kc in 32 64; do
  echo "== TNB_TC2_KC=$kc"
  TNB_TC2_KC=$kc python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullgolden.py -m gpu -x -q 2>&1 | tail -1
  TNB_TC2_KC=$kc python scripts/batch_exp.py 1 0
  TNB_TC2_KC=$kc python scripts/batch_exp.py 8 4
  TNB_TC2_KC=$kc ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gram_tc2_kernel -c 6 --csv --log-file gpurun_out/r02_tc2_kc$kc.csv python scripts/batch_exp.py 1 0 > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/r02_tc2_kc$kc.csv | head -3
done
