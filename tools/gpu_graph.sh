#!/bin/bash
TAG=${1:-gr}
OUT=gpurun_out
mkdir -p $OUT
for z in 1 0 1 0; do PIMCB_ZEROCOPY=$z python tools/latency_ab.py 2>&1 | sed "s/^/zerocopy=$z /"; done | tee $OUT/${TAG}_zerocopy.txt
PIMCB_ZEROCOPY=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k fused_single 2>&1 | tail -3
