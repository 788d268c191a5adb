#!/bin/bash
mkdir -p gpurun_out
for pb in 1 2 3; do
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --per-gpu-batch $pb > gpurun_out/bench_pb$pb.json 2> gpurun_out/bench_pb$pb.err; echo "pb=$pb rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_pb$pb.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['sweep_roofline']['frac'], d['gpu_launches'])"; tail -n 3 gpurun_out/bench_pb$pb.err
done
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "default rc=$?"; cut -c1-2600 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
