#!/bin/bash
# round 2: config 5 through bench.py (16 sequential panels, product graphs within their memory budget) and the other configs' bench lines
mkdir -p gpurun_out
for c in 5 2 3; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-coo-e2e --no-cpu-baseline > gpurun_out/i1_bench_c$c.log 2>&1
  echo "config $c exit $?"; tail -c 600 gpurun_out/i1_bench_c$c.log; echo
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
