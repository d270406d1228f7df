#!/bin/bash
# Short GPU-box visit: parity tests, microbenchmark, bench A/B of the tau-correlation kernels.
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_pytest_gpu.log
./tools/micro/dmma_peak 2>&1 | tee $OUT/${TAG}_dmma_peak.txt
python bench.py --no-cpu-baseline --no-ab --no-pair --steps 200 > $OUT/${TAG}_bench_corr1.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_corr1.json
python bench.py --no-cpu-baseline --no-ab --no-pair --steps 200 --corr-mode 0 --no-e2e > $OUT/${TAG}_bench_corr0.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_corr0.json
tail -3 $OUT/${TAG}_bench.err
