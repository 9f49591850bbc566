#!/bin/bash
# round 2, final evidence run: all GPU tests, smoke, the bench line and its reference arm, then the profiles
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/f_bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/f_ref.log
timeout 600 python tools/quick_bench.py 1 2 3 4 --reps 8 > gpurun_out/f_qb.log 2>&1
timeout 600 python tools/quick_bench.py 5 --panels 16 --reps 3 > gpurun_out/f_qb5.log 2>&1
tail -2 gpurun_out/f_tests.log gpurun_out/f_smoke.log; tail -c 400 gpurun_out/f_bench.log; tail -c 300 gpurun_out/f_ref.log
bash tools/r02_profiles.sh r2z > gpurun_out/r2z_profiles.log 2>&1
tail -3 gpurun_out/r2z_profiles.log
