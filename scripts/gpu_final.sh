#!/bin/bash
# full GPU evidence pass: all -m gpu tests, smoke, default bench, reference arm, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/t_all.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 900 python bench.py --per-gpu-batch 1 > gpurun_out/bench_pb1.json 2> gpurun_out/bench_pb1.err; echo "bench pb1 rc=$?"; cut -c1-300 gpurun_out/bench_pb1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --per-gpu-batch 1 > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gram_tc2_kernel|gram_tc_kernel|project_tc_kernel|cheb_filter_kernel" -c 6 -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --per-gpu-batch 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
