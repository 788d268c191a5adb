#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round.py -m gpu -q --timeout 300 2>&1 | tail -n 4
timeout 900 python scripts/bench_extra.py cfg3 2>&1 | tail -n 2 | cut -c1-600
timeout 900 python scripts/bench_extra.py cfg5 2>&1 | tail -n 2 | cut -c1-600
timeout 1500 python scripts/bench_extra.py cfg4 2>&1 | tail -n 3 | cut -c1-900
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e'])"; tail -n 3 gpurun_out/b.err
