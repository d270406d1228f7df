mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:virial_sym_kernel -s 1 -c 1 -f -o gpurun_out/r01zf_prof_virial python tools/virial_ab.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_sym_kernel -s 1 -c 1 -f -o gpurun_out/r01zf_prof_pair python tools/virial_ab.py > /dev/null 2>&1
ls -la gpurun_out | tail -4
