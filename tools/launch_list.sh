#!/bin/bash
# launch list (gpu__time_duration per launch) of two SpGEMM iterations per config
# usage: tools/launch_list.sh <tag> <config> [...]   (extra quick_bench flags via QB_FLAGS)
tag=$1; shift
mkdir -p gpurun_out
for k in "$@"; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}_c${k}.csv \
      python tools/quick_bench.py $k --reps 2 ${QB_FLAGS:-} > gpurun_out/ncu_${tag}_c${k}.log 2>&1
done
