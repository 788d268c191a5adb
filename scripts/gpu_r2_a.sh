#!/bin/bash
# round 2, call A: full GPU suite (incl. the new full-size goldens) + a default bench run
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -30 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cat gpurun_out/r2a_bench.json | head -c 3000
tail -5 gpurun_out/r2a_bench.err
