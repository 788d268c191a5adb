#!/bin/bash
# bench.py at N = 1, 2, 4, 8 in ONE lease, launched exactly like the driver does (torch.distributed.run, 127.0.0.1)
mkdir -p gpurun_out
NS=${NS:-"1 2 4 8"}
STEPS=${STEPS:-10}
for n in $NS; do
  echo "== N=$n"
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps $STEPS --warmup 3 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err
  fi
  echo "rc=$?"
  python - <<P
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${n}gpu.json'))
    print('N=${n}', 'value', round(d['value'],1), 'ms_per_step', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],2), 'clocks', d['clocks'], 'PB', d['run']['per_gpu_batch_used'])
except Exception as e:
    print('N=${n} no json:', e)
P
  tail -3 gpurun_out/r02_bench_${n}gpu.err
  nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8 | tr '\n' ';'; echo
done
