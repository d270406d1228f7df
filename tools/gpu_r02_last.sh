#!/bin/bash
# Trimmed record visit after the bench default moved to bins of 16 x 512: bench + reference arm + ncu launch list and the
# rho / tau-correlation captures (tests and sanitizers: r02zz / r02final, same library).
TAG=${1:-r02last}
OUT=gpurun_out
mkdir -p $OUT/prof
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -2 $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
KF="regex:rho_|isf_corr|bins_|aos_to|ssf_direct|pair_|virial_"
BARGS="--steps 1 --warmup 3 --no-cpu-baseline --peak-seconds 0.02 --no-ab --no-latency --profile none"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $BARGS --no-pair --no-e2e > $OUT/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rho_lattice_mma -s 4 -c 1 -f -o $OUT/${TAG}_prof_rho_lattice python bench.py $BARGS --no-e2e --no-pair > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:isf_corr -s 4 -c 1 -f -o $OUT/${TAG}_prof_corr python bench.py $BARGS --no-e2e --no-pair > /dev/null 2>&1
python tools/ncu_summary.py $TAG $OUT/prof > /dev/null 2>&1
rm -f $OUT/*.ncu-rep
python -c "
import json
d=json.load(open('$OUT/${TAG}_bench.json')); r=d['roofline']
print('value %.0f e2e %.0f lat %.2f frac %.4f nominal %.4f rho %.2f corr %.2f traffic %s pair %.3f virial %.3f cpu %.5f' % (d['value'], d['e2e']['value'], d['latency']['single_configuration_us'], r['frac'], r['frac_vs_nominal'], r['us_per_64_configurations'], r['corr_kernel']['us_per_64_configurations'], r['traffic'], d['pair_sums']['ms_per_64_configurations'], d['pair_sums']['virial_sums']['ms_per_64_configurations'], d['cpu_baseline']['value']))"
ls $OUT/prof | grep $TAG
