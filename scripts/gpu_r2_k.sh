#!/bin/bash
# driver-like pass: GPU suite, smoke, both bench arms at N=1 (and N=2 when two GPUs are visible)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/r02_bench_final_1gpu.err; echo "bench rc=$?"
python - <<'P'
import json
r=json.load(open('gpurun_out/r02_bench_reference.json')); d=json.load(open('gpurun_out/r02_bench_final_1gpu.json'))
print('reference', r['value'], r['cpu_baseline']['cores'], r['cpu_baseline']['algorithm'])
print('ours', d['value'], 'e2e', d['e2e']['value'], 'single', d['sweep_roofline']['single_call_ms'], 'roof', d['roofline']['frac'], d['roofline']['traffic'], 'clocks', d['clocks'])
print('cpu_baseline', d['cpu_baseline'] and d['cpu_baseline']['value'], 'same_sample', d['same_sample'] and (d['same_sample']['same_config_ratio_device'], d['same_sample']['same_config_ratio_e2e']))
print('ratio device', d['value']/r['value'], 'e2e', d['e2e']['value']/r['value'])
P
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | head -c 300; echo
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | head -c 400; echo
fi
