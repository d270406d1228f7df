#!/bin/bash
# A/B builds of libpimc_b200 with different compile-time switches, for one GPU-box visit:
#   tools/build_variants.sh name1 "-DFOO=1" name2 "-DFOO=2" ...   ->  pimc_b200/variants/libpimc_b200_<name>.so
# Selected at run time with PIMCB_LIB_PATH=<that file> (pimc_b200/build.py).  The .so files are git-ignored.
set -e
cd "$(dirname "$0")/.."
mkdir -p pimc_b200/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --fmad=true \
      -cudart shared $flags -o pimc_b200/variants/libpimc_b200_${name}.so pimc_b200/csrc/pimcb.cu &
done
wait
ls -la pimc_b200/variants
