#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both rho modes), ncu launch list + full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
python bench.py --rho-mode 0 --no-cpu-baseline > $OUT/${TAG}_bench_generic.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_generic.json
KF="regex:rho_|isf_corr|bins_acc|aos_to|ssf_direct|pair_"
BARGS="--steps 8 --warmup 3 --no-cpu-baseline --peak-seconds 0.02 --no-ab --no-pair --no-latency --profile none"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $BARGS > $OUT/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rho_lattice_mma -s 4 -c 1 -f -o $OUT/${TAG}_prof_rho_lattice python bench.py $BARGS --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rho_generic -s 4 -c 1 -f -o $OUT/${TAG}_prof_rho_generic python bench.py $BARGS --no-e2e --rho-mode 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:isf_corr -s 4 -c 1 -f -o $OUT/${TAG}_prof_corr python bench.py $BARGS --no-e2e > /dev/null 2>&1
ls -la $OUT
ncu --set full --clock-control none --import-source on -k regex:pair_sym_kernel -s 2 -c 1 -f -o $OUT/${TAG}_prof_pair python bench.py --steps 4 --warmup 3 --no-cpu-baseline --peak-seconds 0.02 --no-ab --no-e2e > /dev/null 2>&1
ls -la $OUT | tail -12
ncu --set full --clock-control none --import-source on -k regex:virial_sym_kernel -c 1 -f -o $OUT/${TAG}_prof_virial python bench.py --steps 4 --warmup 3 --no-cpu-baseline --peak-seconds 0.02 --no-ab --no-e2e > /dev/null 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_ref.json
ls -la $OUT | tail -14
