#!/bin/bash
# GPU-box visit for a round's new work: the new parity tests first (fail fast, logged), then the whole GPU suite with
# durations, smoke, the default bench, the reference arm, and the ncu launch list.  Usage: bash tools/gpu_check.sh <tag>
TAG=${1:-chk}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt
python -m pytest tests/test_variants.py tests/test_comm.py tests/test_reference_gpu.py -m gpu -q 2>&1 | tail -40 > $OUT/${TAG}_pytest_new.log; cat $OUT/${TAG}_pytest_new.log
python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -60 > $OUT/${TAG}_pytest_gpu.log; tail -30 $OUT/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_ref.json
KF="regex:rho_|isf_corr|bins_|aos_to|ssf_|pair_|virial_|elastic_"
BARGS="--steps 8 --warmup 3 --no-cpu-baseline --peak-seconds 0.02 --no-ab --no-latency --profile none"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $BARGS > $OUT/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:virial_kernel -c 1 -f -o $OUT/${TAG}_prof_virial python bench.py $BARGS --no-e2e > /dev/null 2>&1
ls -la $OUT | tail -15
