#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tucker.py tests/test_gpu_cp.py tests/test_gpu_round.py -m gpu -q --timeout 300 > gpurun_out/t_tk.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/t_tk.log | cut -c1-220
timeout 1500 python scripts/bench_extra.py cfg4 2>&1 | tail -n 2 | cut -c1-700
