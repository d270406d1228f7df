#!/bin/bash
# ncu --set full of the persistent tau-correlation kernel (corr mode 2); only CSV exports travel back.
OUT=gpurun_out
BARGS="--steps 2 --warmup 3 --no-cpu-baseline --peak-seconds 0.02 --no-ab --no-latency --profile none --no-e2e --no-pair"
PIMCB_CORR_MODE=2 ncu --set full --clock-control none --import-source on -k regex:isf_corr_mma_pipe -s 4 -c 1 -f -o $OUT/corrpipe python bench.py $BARGS > /dev/null 2>&1
ncu -i $OUT/corrpipe.ncu-rep --page source --csv --print-source sass > $OUT/corrpipe_source.csv 2>/dev/null
ncu -i $OUT/corrpipe.ncu-rep --page raw --csv > $OUT/corrpipe_raw.csv 2>/dev/null
python tools/ncu_hot.py $OUT/corrpipe.ncu-rep 0.008 > $OUT/corrpipe_hot.txt 2>&1
rm -f $OUT/corrpipe.ncu-rep
ls -la $OUT | grep corrpipe
