#!/bin/bash
# round 2, product graphs: parity subset, per-config times with and without graphs, the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "graphs or size_plans or owner_variants or panels or fuzz or pool_reuse" > gpurun_out/g1_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/g1_tests.log
for g in 0 1; do
  timeout 300 python tools/quick_bench.py 1 2 4 --reps 8 --graphs $g > gpurun_out/g1_qb_graphs$g.log 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-coo-e2e --no-cpu-baseline > gpurun_out/g1_bench.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --config 1 --no-coo-e2e --no-cpu-baseline > gpurun_out/g1_bench_c1.log 2>&1
tail -3 gpurun_out/g1_tests.log; tail -4 gpurun_out/g1_qb_graphs*.log; tail -c 1500 gpurun_out/g1_bench.log
