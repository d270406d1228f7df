#!/bin/bash
# tau-correlation with persistent CTAs (corr mode 2) against mode 1: parity tests, then C2 at 64 and 256 configurations per launch.
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tau_correlation or odd_slice" 2>&1 | tail -5
for b in 64 256; do
  for mode in 1 2; do
    PIMCB_CORR_MODE=$mode python bench.py --batch $b --batches-per-step $((1024 / b)) --steps 20 --warmup 5 --no-cpu-baseline --no-ab --no-pair --no-latency --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('batch $b corr_mode $mode value %.0f rho_us_per_64 %.2f corr_us_per_64 %.2f' % (d['value'], r['us_per_64_configurations'], r['corr_kernel']['us_per_64_configurations']))"
  done
done
