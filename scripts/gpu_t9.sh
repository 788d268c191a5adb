#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cp.py tests/test_gpu_round.py tests/test_gpu_tucker.py tests/test_gpu_blocks.py -m gpu -q --timeout 300 2>&1 | tail -n 6 | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --per-gpu-batch 1 --no-e2e --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err; tail -n 3 gpurun_out/b1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b1.json'))
print('single in flight', d['value'], d['ms_per_step'], json.dumps(d.get("phases_ms", {}))[:900])
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-concurrent-flag > gpurun_out/b4.json 2> gpurun_out/b4.err; tail -n 3 gpurun_out/b4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b4.json'))
print('4 in flight', d['value'], d['ms_per_step'])
PY
