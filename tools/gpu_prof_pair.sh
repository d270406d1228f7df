#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
export PIMCB_TABLE_CODEC=${CODEC:-0}
ncu --set full --clock-control none --import-source on -k regex:"pair_tile|virial_tile" --launch-skip 1 --launch-skip-before-match 0 -c 5 -f -o $OUT/${TAG}_pair_tile python tools/pair_prof.py > $OUT/${TAG}_ncu.log 2>&1
tail -5 $OUT/${TAG}_ncu.log
ncu -i $OUT/${TAG}_pair_tile.ncu-rep --page raw --csv > $OUT/${TAG}_pair_tile_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_pair_tile.ncu-rep --page source --csv > $OUT/${TAG}_pair_tile_source.csv 2>/dev/null
rm -f $OUT/${TAG}_pair_tile.ncu-rep      # gpurun brings back at most 64 MiB: the CSV pages carry what the summaries need
ls -la $OUT | tail -5
