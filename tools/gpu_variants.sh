#!/bin/bash
# Pair / virial probe over the compiled variants in pimc_b200/variants (tools/build_variants.sh)
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
export SUBS=1
echo "== default build" | tee $OUT/${TAG}_variants.txt
python tools/pair_probe.py 2>&1 | tee -a $OUT/${TAG}_variants.txt
for f in pimc_b200/variants/libpimc_b200_*.so; do
    echo "== $f" | tee -a $OUT/${TAG}_variants.txt
    PIMCB_LIB_PATH=$PWD/$f python tools/pair_probe.py 2>&1 | tee -a $OUT/${TAG}_variants.txt
done
