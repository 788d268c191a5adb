#!/bin/bash
mkdir -p gpurun_out
# (a CP-ALS debug print ran here; the script was removed with the other scratch helpers)
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
P
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python /tmp/one.py > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches.csv 30 > gpurun_out/r02_launch_summary.txt
: > gpurun_out/r02_ncu_summaries_raw.md
for k in gram_tc2_kernel project_tc_kernel gram_tc_kernel cheb_filter_kernel cd_rr_kernel cd_chol_kernel jacobi2_eigh_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^${k} -c 2 -f -o /tmp/r02_${k} python /tmp/one.py > /dev/null 2>&1
  echo "## ${k}" >> gpurun_out/r02_ncu_summaries_raw.md
  python scripts/ncu_summarize.py /tmp/r02_${k}.ncu-rep >> gpurun_out/r02_ncu_summaries_raw.md 2>&1
done
# keep the two dominant reports (one launch each) for the record
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^gram_tc2_kernel -c 1 -f -o gpurun_out/r02_gram_tc2_kernel python /tmp/one.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^cheb_filter_kernel -c 1 -s 1 -f -o gpurun_out/r02_cheb_filter_kernel python /tmp/one.py > /dev/null 2>&1
ls -la gpurun_out | awk '{print $5, $9}'
wc -l gpurun_out/r02_ncu_summaries_raw.md
