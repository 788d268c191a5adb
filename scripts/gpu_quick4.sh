#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 120 > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; tail -n 12 gpurun_out/t_tc.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_ttsvd.py tests/test_gpu_round.py -m gpu -q -x --timeout 300 > gpurun_out/t_ttsvd.log 2>&1; echo "ttsvd rc=$?"; tail -n 4 gpurun_out/t_ttsvd.log | cut -c1-200
for pb in 1 4; do
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --per-gpu-batch $pb > gpurun_out/bench_pb$pb.json 2> gpurun_out/bench_pb$pb.err; echo "pb=$pb rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_pb$pb.json')); print(d['value'], d['ms_per_step'], d['phases_ms'], d['rel_error'])"; tail -n 3 gpurun_out/bench_pb$pb.err
done
