#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cp.py tests/test_gpu_blocks.py -m gpu -q --timeout 300 > gpurun_out/t_cp.log 2>&1; echo "cp rc=$?"; tail -n 30 gpurun_out/t_cp.log | cut -c1-300
timeout 300 python scripts/gpu_diag.py blocks 2>&1 | grep jacobi
