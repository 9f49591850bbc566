#!/bin/bash
# round 2: step-3 record format + window-kernel loops: parity subset and per-config times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "graphs or size_plans or fuzz or dense" > gpurun_out/h1_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/h1_tests.log
timeout 400 python tools/quick_bench.py 1 4 --reps 5 > gpurun_out/h1_qb.log 2>&1
tail -3 gpurun_out/h1_tests.log; grep -v "^   rep" gpurun_out/h1_qb.log | tail -12
