#!/bin/bash
# round 2 evidence pass on ONE B200: tests, smoke, TF32 peak, launch list, ncu --set full captures of the dominant kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -14 gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/measure_tf32_peak.py gpurun_out/r02_tf32_peak.json 2>&1 | tail -1
cat > /tmp/one.py <<'P'
import torch, sys
sys.path.insert(0, '.')
from tntorch_b200 import ops
g = torch.Generator(device="cuda").manual_seed(7)
X = torch.randn((64,) * 5, generator=g, device="cuda")
plan = ops.TTSVDPlan((64,)*5, torch.float32, rmax=32)
for _ in range(2): plan.run(X)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
plan.run(X)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('speculative', plan.info[26], 'products', plan.info[2], 'rr steps', plan.info[30])
P
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python /tmp/one.py > gpurun_out/r02_one.log 2>&1
tail -1 gpurun_out/r02_one.log
python scripts/launch_summary.py gpurun_out/r02_launches.csv 30 | tee gpurun_out/r02_launch_summary.txt
for k in gram_tc2_kernel project_tc_kernel gram_tc_kernel cheb_filter_kernel cd_rr_kernel jacobi2_eigh_kernel cd_chol_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^${k} -c 2 -f -o gpurun_out/r02_${k} python /tmp/one.py > gpurun_out/r02_ncu_${k}.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
python scripts/bench_extra.py cfg3 cfg5 2>&1 | tail -1 | tee gpurun_out/r02_bench_extra.json
