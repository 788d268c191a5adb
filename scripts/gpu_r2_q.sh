#!/bin/bash
# round 2, pass Q: cfg4 per-sweep time from 21-sweep and 1-sweep runs (the 1.1 s initialisation dominates a 5-sweep difference)
mkdir -p gpurun_out
python scripts/bench_extra.py cfg4 2>&1 | tail -1 | tee gpurun_out/r02_cfg4.json
