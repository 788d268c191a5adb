#!/bin/bash
# round 2, pass O (2 GPUs): torchrun sanity of both bench arms at N=2 after the kernel changes, and an `ncu --set full`
# look at the Khatri-Rao reduction of the CP-ALS sweep (one GPU)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | head -c 400; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err; echo "bench2 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu_final.json'))
print('N=2', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], d['clocks'])
P
cat > /tmp/cp2.py <<'P'
import sys, torch
sys.path.insert(0, ".")
from tntorch_b200 import ops
X = torch.randn(256, 256, 256, 256, device="cuda")
ops.cp_als(X, 50, max_iter=1, tol=float("-inf"))
torch.cuda.synchronize()
P
ncu --set full --clock-control none -k regex:khatri_reduce_kernel -c 7 -f -o /tmp/khatri python /tmp/cp2.py > /dev/null 2>&1
python scripts/ncu_summarize.py /tmp/khatri.ncu-rep > gpurun_out/r02_khatri_ncu.md 2>&1
grep -E "Kernel Name|Grid Size|gpu__time_duration|dram__bytes_read|gpu__dram_throughput|sm__warps_active|sm__issue_active|registers|lts__t_sector_hit|fp64" gpurun_out/r02_khatri_ncu.md | head -80
