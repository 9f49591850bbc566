#!/bin/bash
# Does the reference rebuilt with its OWN virtual architecture (compute_61 PTX -> sm_100 SASS) get through
# its NSPARSE step 1 on a B200?  lap600 first (smallest NSPARSE-path input); the BASELINE configs only if it does.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/ref_run.py --bin pemspgemm_ref61 --timeout 90 lap600 > gpurun_out/ref61_lap600.log 2>&1
cat gpurun_out/ref61_lap600.log
if grep -q "rc 0" gpurun_out/ref61_lap600.log; then
  python tools/ref_run.py --bin pemspgemm_ref61 --timeout 240 1 2 3 4 > gpurun_out/ref61_configs.log 2>&1
  cat gpurun_out/ref61_configs.log
  python tests/golden/make_golden.py --bin pemspgemm_ref61 --timeout 120 rand300k_a2 lap600_a2 webbase270k_a2 > gpurun_out/golden_ref61.log 2>&1
  tail -20 gpurun_out/golden_ref61.log
fi
