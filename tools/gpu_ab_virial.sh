mkdir -p gpurun_out
for v in zz0 zz1 zz0 zz1; do PIMCB_LIB_PATH=$PWD/pimc_b200/variants/libpimc_b200_$v.so python tools/virial_ab.py 2>&1 | tail -1; done | tee gpurun_out/r01zz_ring_ab.txt
python -m pytest tests/test_gpu_parity.py tests/test_variants.py tests/test_reference_cpu.py tests/test_golden.py -m gpu -q -k "pair or virial or upstream or golden" 2>&1 | tail -5 | tee gpurun_out/r01zz_pytest.log
