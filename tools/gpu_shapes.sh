#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_gpu.py tests/test_golden_upstream.py -m gpu -x -q 2>&1 | tail -3
for w in C3 C2 C4; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-ab --no-pair --no-latency --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$w value %.0f rho_ms %.4f corr_ms %.4f frac %.4f tiles %dx%d' % (d['value'], r['avg_launch_ms'], r['corr_kernel']['avg_launch_ms'], r['frac'], r['plan']['M_tiles'], r['plan']['N_tiles']))"
done
