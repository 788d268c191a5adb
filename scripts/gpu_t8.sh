#!/bin/bash
mkdir -p gpurun_out
sed -n '/^cat > \/tmp\/proj_only.py/,/^PY$/p' scripts/gpu_ncu_proj.sh | sed '1d;$d' > /tmp/proj_only.py
timeout 300 python /tmp/proj_only.py
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ttsvd.py -m gpu -q --timeout 300 2>&1 | tail -n 6 | cut -c1-300
