#!/bin/bash
# Two-GPU visit: the library's own NCCL exchange step from Python (tests) and from the C++ CLI (pimcb_measure --batch --nranks 2).
TAG=${1:-g2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | tee $OUT/${TAG}_gpu.txt
python -m pytest tests/test_comm.py tests/test_host_layer.py -m gpu -q -k "comm or rank or batched" 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
