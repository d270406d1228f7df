for b in 256 512 1024; do
python bench.py --batch $b --batches-per-step $((1024 / b)) --steps 20 --warmup 5 --no-cpu-baseline --no-ab --no-pair --no-latency --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('batch $b value %.0f rho_us_per_64 %.2f corr_us_per_64 %.2f frac %.4f' % (d['value'], r['us_per_64_configurations'], r['corr_kernel']['us_per_64_configurations'], r['frac']))"
done
