#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_variants.py tests/test_reference_cpu.py tests/test_host_layer.py -m gpu -x -q -k "pair or potential or action or energy or virial or force or graph_replay_leaves" 2>&1 | tail -4
for sp in 1 0; do
  PIMCB_PAIR_SPLIT=$sp PIMCB_VIRIAL_SPLIT=$sp SUBS="1" python tools/pair_probe.py 2>&1 | tail -1
done
