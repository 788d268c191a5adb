#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_ttsvd.py -m gpu -q -x --timeout 300 > gpurun_out/t_blocks.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_blocks.log
timeout 600 python scripts/gpu_diag.py blocks 2>&1 | grep jacobi
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['phases_ms'], d['gpu_launches'])"; tail -n 5 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:project_f32_kernel -c 1 -o gpurun_out/prof_project python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_proj.log 2>&1; echo "ncu rc=$?"
