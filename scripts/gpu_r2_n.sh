#!/bin/bash
# round 2, final evidence pass N (after the gram_tc2 stage change): all -m gpu tests, smoke, default bench, launch list,
# `ncu --set full` summaries of the main kernels and the DRAM traffic file bench.py reads
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
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
python scripts/launch_summary.py gpurun_out/r02_launches.csv 30 > gpurun_out/r02_launch_summary.txt; head -8 gpurun_out/r02_launch_summary.txt
: > gpurun_out/r02_ncu_summaries_raw.md
for k in gram_tc2_kernel project_tc_kernel gram_tc_kernel cheb_filter_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^${k} -c 2 -f -o /tmp/r02_${k} python /tmp/one.py > /dev/null 2>&1
  echo "## ${k}" >> gpurun_out/r02_ncu_summaries_raw.md
  python scripts/ncu_summarize.py /tmp/r02_${k}.ncu-rep >> gpurun_out/r02_ncu_summaries_raw.md 2>&1
done
cp /tmp/r02_gram_tc2_kernel.ncu-rep gpurun_out/r02_gram_tc2_kernel.ncu-rep
python scripts/traffic_from_summaries.py gpurun_out/r02_ncu_summaries_raw.md profiles/r02_ncu_traffic.json && cp profiles/r02_ncu_traffic.json gpurun_out/r02_ncu_traffic.json
timeout 900 python bench.py --gpus 1 > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_final_1gpu.json'))
print('ours', d['value'], 'e2e', d['e2e']['value'], 'single', d['sweep_roofline']['single_call_ms'], 'roof', d['roofline']['frac'], d['roofline']['traffic'], 'sweep frac', d['sweep_roofline']['frac'], 'clocks', d['clocks'], 'phases', d['phases_ms'])
r=json.load(open('gpurun_out/r02_bench_reference.json')); print('reference', r['value'])
P
