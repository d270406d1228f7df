#!/bin/bash
TAG=${1:-gr}
OUT=gpurun_out
mkdir -p $OUT
python tools/latency_breakdown.py 2>&1 | tee $OUT/${TAG}_latency_breakdown.txt
