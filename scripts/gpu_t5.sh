#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 120 2>&1 | tail -n 4 | cut -c1-300
for extra in "" "--no-concurrent-flag"; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $extra > gpurun_out/b4.json 2> gpurun_out/b4.err; tail -n 3 gpurun_out/b4.err
python - "$extra" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b4.json'))
print('4 in flight', sys.argv[1], d['value'], d['ms_per_step'])
PY
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --per-gpu-batch 6 > gpurun_out/b6.json 2> gpurun_out/b6.err; tail -n 3 gpurun_out/b6.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b6.json'))
print('6 in flight', d['value'], d['ms_per_step'])
PY
