#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 60 > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; tail -n 15 gpurun_out/t_tc.log | cut -c1-220
timeout 300 python scripts/gpu_diag.py tc 2>&1 | tail -n 7
TNB_GRAM_TC2=0 timeout 300 python scripts/gpu_diag.py tc 2>&1 | grep "2048)"
