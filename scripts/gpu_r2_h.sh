#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spec.py tests/test_gpu_blocks.py tests/test_gpu_round.py -x -q 2>&1 | tail -5
NS="1" STEPS=10 bash scripts/gpu_r2_scale.sh
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
for k in ('value','ms_per_step','run','rel_error','rel_error_twin','gpu_launches','phases_ms','sweep_roofline','e2e','roofline'): print(k, d.get(k))
P
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | head -c 1500
