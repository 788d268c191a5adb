#!/bin/bash
# A/B of the cta_group::2 Gram kernel's pipeline depth: TNB_TC2_KC=32 (4 KB boxes x 6 stages) vs 64 (8 KB boxes x 3 stages)
mkdir -p gpurun_out
for kc in 32 64; do
  echo "== TNB_TC2_KC=$kc"
  TNB_TC2_KC=$kc python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullgolden.py -m gpu -x -q 2>&1 | tail -1
  TNB_TC2_KC=$kc python scripts/batch_exp.py 1 0
  TNB_TC2_KC=$kc python scripts/batch_exp.py 8 4
  TNB_TC2_KC=$kc ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gram_tc2_kernel -c 6 --csv --log-file gpurun_out/r02_tc2_kc$kc.csv python scripts/batch_exp.py 1 0 > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/r02_tc2_kc$kc.csv | head -3
done
