#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round.py tests/test_gpu_tucker.py tests/test_gpu_blocks.py tests/test_gpu_ttsvd.py -m gpu -q --timeout 300 > gpurun_out/t3.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/t3.log | cut -c1-220
timeout 600 python scripts/gpu_diag.py blocks 2>&1 | grep jacobi
timeout 900 python scripts/bench_extra.py cfg3 2>&1 | tail -n 1 | cut -c1-400
