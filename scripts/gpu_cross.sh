#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cross.py -m gpu -q --timeout 300 > gpurun_out/t_cross.log 2>&1; echo "cross rc=$?"; tail -n 40 gpurun_out/t_cross.log | cut -c1-250
