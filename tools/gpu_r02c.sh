#!/bin/bash
# Round-2 record visit (one B200): GPU test suite, smoke, bench (C2 + the other workloads + reference arm), ncu launch
# list and full captures summarised on the box (only the summaries travel back), sanitizers.
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT/prof
nvidia-smi -L > $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -2 $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
for W in C1 C3 C4; do
    timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-ab > $OUT/${TAG}_bench_${W}.json 2>> $OUT/${TAG}_bench.err
done
KF="regex:rho_|isf_corr|bins_|aos_to|ssf_direct|pair_|virial_"
BARGS="--steps 2 --warmup 3 --no-cpu-baseline --peak-seconds 0.02 --no-ab --no-latency --profile none"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $BARGS > $OUT/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rho_lattice_mma -s 4 -c 1 -f -o $OUT/${TAG}_prof_rho_lattice python bench.py $BARGS --no-e2e --no-pair > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:isf_corr -s 4 -c 1 -f -o $OUT/${TAG}_prof_corr python bench.py $BARGS --no-e2e --no-pair > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fp64_peak -s 2 -c 1 -f -o $OUT/${TAG}_prof_fp64_peak python bench.py $BARGS --no-e2e --no-pair > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_tile_kernel -s 2 -c 1 -f -o $OUT/${TAG}_prof_pair_tile python bench.py $BARGS --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:virial_tile_kernel -s 1 -c 1 -f -o $OUT/${TAG}_prof_virial_tile python bench.py $BARGS --no-e2e > /dev/null 2>&1
python tools/ncu_summary.py $TAG $OUT/prof > /dev/null 2>&1
rm -f $OUT/*.ncu-rep
ls -la $OUT $OUT/prof | tail -30
bash tools/gpu_sanitize.sh ${TAG}
