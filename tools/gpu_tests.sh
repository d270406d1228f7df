#!/bin/bash
# GPU-box visit: whole GPU suite (durations), smoke.  Usage: bash tools/gpu_tests.sh <tag>
TAG=${1:-t}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > $OUT/${TAG}_pytest_gpu.log; tail -25 $OUT/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.log
