#!/bin/bash
# GPU visit: tests of the pair / virial paths, then the pair probe in the three table modes.
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_variants.py tests/test_reference_cpu.py tests/test_golden_upstream.py tests/test_golden.py tests/test_host_layer.py -m gpu -x -q -k "pair or virial or force or upstream or golden or external or potential or energy" 2>&1 | tail -15 > $OUT/${TAG}_pytest_pair.log; cat $OUT/${TAG}_pytest_pair.log
echo "== default (packed when force slices)"; timeout 300 python tools/pair_probe.py 2>&1 | tee $OUT/${TAG}_probe_default.txt
echo "== packed everywhere"; PIMCB_PAIR_PACKED=1 timeout 300 python tools/pair_probe.py 2>&1 | tee $OUT/${TAG}_probe_packed.txt
echo "== verbatim everywhere"; PIMCB_TABLE_CODEC=0 timeout 300 python tools/pair_probe.py 2>&1 | tee $OUT/${TAG}_probe_verbatim.txt
