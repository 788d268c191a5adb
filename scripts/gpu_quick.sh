#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_tc.py -m gpu -q -x --timeout 300 > gpurun_out/t_blocks.log 2>&1; echo "blocks rc=$?"; tail -n 4 gpurun_out/t_blocks.log
timeout 900 python -m pytest tests/test_gpu_ttsvd.py tests/test_gpu_round.py -m gpu -q --timeout 300 > gpurun_out/t_ttsvd.log 2>&1; echo "ttsvd rc=$?"; tail -n 6 gpurun_out/t_ttsvd.log
timeout 600 python scripts/gpu_diag.py blocks > gpurun_out/d_blocks.log 2>&1; cat gpurun_out/d_blocks.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-2500; tail -n 5 gpurun_out/bench.err
