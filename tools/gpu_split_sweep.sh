#!/bin/bash
# rho kernel: uniform slice split (PIMCB_RHO_SPLIT) at 64 and 256 configurations per launch, C2, device-resident.
OUT=gpurun_out
for b in 64 256; do
  for sp in 0 1 2 4 8; do
    if [ $sp = 0 ]; then unset PIMCB_RHO_SPLIT; else export PIMCB_RHO_SPLIT=$sp; fi
    python bench.py --batch $b --batches-per-step $((1024 / b)) --steps 20 --warmup 5 --no-cpu-baseline --no-ab --no-pair --no-latency --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('batch $b split $sp value %.0f rho_us_per_64 %.2f corr_us_per_64 %.2f frac %.4f' % (d['value'], r['us_per_64_configurations'], r['corr_kernel']['us_per_64_configurations'], r['frac']))"
  done
done
