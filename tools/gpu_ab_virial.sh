mkdir -p gpurun_out
for v in vu1 vu2 vu3 vu4 vu2b3 vu3b3; do PIMCB_LIB_PATH=$PWD/pimc_b200/variants/libpimc_b200_$v.so python tools/virial_ab.py 2>&1 | tail -1; done | tee gpurun_out/r01y_virial_ab.txt
python -m pytest tests/test_comm.py tests/test_variants.py -q 2>&1 | tail -5 | tee gpurun_out/r01y_pytest_comm.log
