#!/bin/bash
# The single-walker graph with the rho kernel reading the beads array directly (PIMCB_RHO_DIRECT=1, default) vs the
# transpose node in front (=0): parity tests, then latency and device-resident value for C2 / C1 / C3 / C4.
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph or fused or split or lattice or fuzz" 2>&1 | tail -4
PIMCB_RHO_DIRECT=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph or fused" 2>&1 | tail -2
for w in C2 C1 C3 C4; do
  for d in 1 0; do
    PIMCB_RHO_DIRECT=$d python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-ab --no-pair 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$w direct $d value %.0f latency_us %.2f rho_us_per_64 %.2f frac %.4f e2e %.0f' % (d['value'], d['latency']['single_configuration_us'], r['us_per_64_configurations'], r['frac'], d['e2e']['value']))"
  done
done
