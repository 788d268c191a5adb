#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arith.py tests/test_gpu_cross.py tests/test_gpu_round.py tests/test_gpu_tc.py tests/test_gpu_cp.py tests/test_gpu_spec.py tests/test_gpu_batch.py -x -q -s 2>&1 | tail -25 | tee gpurun_out/r2f_tests.log
for order in wave phase; do for gate in 0 1; do
  echo "== order=$order no_gate=$gate"
  if [ $gate = 1 ]; then export TNB_NO_GATE=1; else unset TNB_NO_GATE; fi
  TNB_BATCH_ORDER=$order timeout 300 python scripts/batch_exp.py 6 4 2>&1 | tail -1
done; done
unset TNB_NO_GATE
for b in 3 4 8; do timeout 300 python scripts/batch_exp.py $b 4 2>&1 | tail -1; done
timeout 300 python scripts/batch_exp.py 6 0 2>&1 | tail -1
TNB_NO_GATE=1 timeout 300 python scripts/batch_exp.py 6 0 2>&1 | tail -1
timeout 300 python scripts/batch_exp.py 1 0 2>&1 | tail -1
timeout 120 python scripts/measure_tf32_peak.py 2>&1 | tail -2
