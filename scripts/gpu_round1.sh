#!/bin/bash
# first GPU pass: build check, block tests, tensor-core tests (own process + timeout), parity tests, timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import tntorch_b200; from tntorch_b200 import ops; print('lib', ops.lib().tnb_version())" > gpurun_out/import.log 2>&1
timeout 900 python -m pytest tests/test_gpu_blocks.py -m gpu -q -x --timeout 300 > gpurun_out/t_blocks.log 2>&1; echo "blocks rc=$?"
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 120 > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"
timeout 600 python scripts/gpu_diag.py tc > gpurun_out/d_tc.log 2>&1; echo "diag tc rc=$?"
timeout 900 python -m pytest tests/test_gpu_ttsvd.py tests/test_gpu_round.py -m gpu -q --timeout 300 > gpurun_out/t_ttsvd.log 2>&1; echo "ttsvd rc=$?"
timeout 600 python scripts/gpu_diag.py blocks > gpurun_out/d_blocks.log 2>&1
timeout 900 python scripts/gpu_diag.py ttsvd > gpurun_out/d_ttsvd.log 2>&1; echo "diag ttsvd rc=$?"
timeout 900 python scripts/gpu_diag.py big > gpurun_out/d_big.log 2>&1; echo "diag big rc=$?"
for f in t_blocks t_tc t_ttsvd; do tail -n 5 gpurun_out/$f.log; done; tail -n 20 gpurun_out/d_tc.log; tail -n 20 gpurun_out/d_big.log
