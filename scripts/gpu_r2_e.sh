#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_batch.py -x -q -s 2>&1 | tail -30 | tee gpurun_out/r2e_spec.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_spec.py --deselect tests/test_gpu_batch.py 2>&1 | tail -15 | tee gpurun_out/r2e_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
for k in ('value','ms_per_step','run','rel_error','rel_error_twin','gpu_launches','phases_ms','sweep_roofline','e2e','same_sample'): print(k, d.get(k))
print(d['roofline'])
P
tail -5 gpurun_out/r2e_bench.err
