#!/bin/bash
# Round-1 profiling pass (run on the GPU box through gpurun):
#   1. launch list (gpu__time_duration per launch) of two SpGEMM iterations per config
#   2. one `ncu --set full` capture of the hot kernels of the second iteration, per config
# usage: tools/ncu_capture.sh <tag> <config> [<config> ...]
set -u
tag=$1; shift
mkdir -p gpurun_out
for k in "$@"; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}_c${k}.csv \
      python tools/quick_bench.py $k --reps 2 > gpurun_out/ncu_${tag}_c${k}.log 2>&1
  ncu --set full --clock-control none --import-source on \
      -k regex:'k_expand|k_step2_masks_tile|k_step3_entries|k_rowcolidx|k_compact|k_step2_|k_step3_' --launch-skip ${SKIP:-5} -c ${COUNT:-5} \
      -f -o gpurun_out/full_${tag}_c${k} python tools/quick_bench.py $k --reps 2 >> gpurun_out/ncu_${tag}_c${k}.log 2>&1
done
ls -la gpurun_out | tail -20
