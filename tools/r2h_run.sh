#!/bin/bash
# developer loop on the GPU box: a parity subset, then per-config times (usage: tools/r2h_run.sh ["pytest -k expr"] [configs])
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "${1:-graph or dense or fuzz or owner_variants}" > gpurun_out/h1_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/h1_tests.log
timeout 400 python tools/quick_bench.py ${2:-1 4} --reps 6 --check > gpurun_out/h1_qb.log 2>&1
tail -3 gpurun_out/h1_tests.log; grep -v "^   rep" gpurun_out/h1_qb.log | tail -12
