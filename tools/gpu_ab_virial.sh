mkdir -p gpurun_out
for v in 1 0 1 0; do PIMCB_PAIR_SYM=$v python tools/virial_ab.py 2>&1 | tail -1 | sed "s/^/pair_sym=$v /"; done | tee gpurun_out/r01ps_pair_sym_ab.txt
python -m pytest tests -m gpu -q -k "pair or golden or upstream or energy or plugin or virial or external or smoke" 2>&1 | tail -5 | tee gpurun_out/r01ps_pytest.log
