#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 env "$@" > gpurun_out/bx.json 2> gpurun_out/bx.err; tail -n 2 gpurun_out/bx.err; python - "$tag" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bx.json')); print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), d['config'].get('reserved_sms'), d.get('rel_error'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
}
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
run "pb4 conc R4 filter" X=1 $B
run "pb4 conc R4 nofilter" TNB_NO_RESIDENT_FILTER=1 $B
run "pb6 conc R4 filter" X=1 $B --per-gpu-batch 6
run "pb6 conc R4 nofilter" TNB_NO_RESIDENT_FILTER=1 $B --per-gpu-batch 6
run "pb8 conc R4 nofilter" TNB_NO_RESIDENT_FILTER=1 $B --per-gpu-batch 8
run "pb1 filter" X=1 $B --per-gpu-batch 1
run "pb1 nofilter" TNB_NO_RESIDENT_FILTER=1 $B --per-gpu-batch 1
