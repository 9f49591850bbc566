#!/bin/bash
# one `ncu --set full` capture of selected kernels in the second SpGEMM iteration
# usage: tools/ncu_full.sh <tag> <config> <kernel-regex> <skip> <count>   (QB_FLAGS for quick_bench flags)
tag=$1; k=$2; rx=$3; skip=$4; cnt=$5
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip -c $cnt \
    -f -o gpurun_out/full_${tag}_c${k} python tools/quick_bench.py $k --reps 2 ${QB_FLAGS:-} > gpurun_out/ncufull_${tag}_c${k}.log 2>&1
