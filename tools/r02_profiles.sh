#!/bin/bash
# Round-2 evidence run (on the GPU box): launch lists of every config and `ncu --set full` of the hot kernels of
# the second SpGEMM iteration, exported to text on the box (the .ncu-rep files are too large to travel back).
cd "$(dirname "$0")/.."
tag=${1:-r2p}
RX='k_step3_entries|k_step3_windows|k_step2_pairs|k_expand|Onesweep|k_row_sort|k_row_tiles|k_tile_products|k_step1_count|k_step1_fill|k_build_tiles|k_ctiles'
full() {   # config, launch-skip, count, extra quick_bench flags
  local k=$1 skip=$2 cnt=$3; shift 3
  timeout 1200 ncu --set full --clock-control none -k regex:"$RX" --launch-skip $skip -c $cnt -f -o gpurun_out/full_${tag}_c$k \
      python tools/quick_bench.py $k "$@" > gpurun_out/ncufull_${tag}_c$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/full_${tag}_c$k.ncu-rep > gpurun_out/ncu_${tag}_c$k.txt 2>&1
  ncu -i gpurun_out/full_${tag}_c$k.ncu-rep --page raw --csv > gpurun_out/ncuraw_${tag}_c$k.csv 2>/dev/null
  rm -f gpurun_out/full_${tag}_c$k.ncu-rep
  tail -n 2 gpurun_out/ncufull_${tag}_c$k.log
}
# (--graphs 0: the kernels of a product are launched one by one, so that --launch-skip counts mean what they say)
for k in 1 2 3 4; do QB_FLAGS="--graphs 0" bash tools/launch_list.sh $tag $k; done
QB_FLAGS="--panels 16 --reps 1 --graphs 0" bash tools/launch_list.sh $tag 5
# kernels matched per product: bitmap path 4 (count, fill, pairs, numeric); radix path 9 (tile_products, expand, 4 x onesweep,
# ctiles, pairs, entries); row-sort path 6 (tile_products, expand, row_sort, row_tiles, pairs, windows); + k_build_tiles per conversion
full 1 5 4 --reps 2 --graphs 0
full 2 10 9 --reps 2 --graphs 0
full 3 12 9 --reps 2 --graphs 0
full 4 7 6 --reps 2 --graphs 0
full 5 10 9 --panels 16 --reps 1 --graphs 0     # second panel of the first product
du -sh gpurun_out
