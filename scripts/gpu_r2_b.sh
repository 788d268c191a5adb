#!/bin/bash
# round 2, call B: new eigen stage — targeted tests first, then the whole suite, then single-call timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py -x -q 2>&1 | tail -40 | tee gpurun_out/r2b_spec.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_spec.py 2>&1 | tail -15 | tee gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --per-gpu-batch 1 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err; echo "bench1 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2b_bench1.json'))
print('single in flight:', d['value'], 'ms', d['ms_per_step'], d['phases_ms'], 'launches', d['gpu_launches'])
P
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench6.json 2> gpurun_out/r2b_bench6.err; echo "bench6 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2b_bench6.json'))
print('six in flight:', d['value'], 'ms', d['ms_per_step'])
P
tail -3 gpurun_out/r2b_bench1.err gpurun_out/r2b_bench6.err
