#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 0" "4 8" "4 16" "6 12"; do set -- $cfg
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --per-gpu-batch $1 --reserve-sms $2 > gpurun_out/b.json 2> gpurun_out/b.err; echo "pb=$1 reserve=$2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['ms_per_step'])"; tail -n 3 gpurun_out/b.err
done
