#!/bin/bash
# Final N-GPU check of the default bench (the command the driver's scaling run uses) with the round's last code.
N=${1:-8}
TAG=${2:-r02final}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline --no-ab --no-pair --no-latency > $OUT/${TAG}_c2_n1.json 2> $OUT/${TAG}_c2_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 > $OUT/${TAG}_c2_n$N.json 2> $OUT/${TAG}_c2_n$N.err
for f in c2_n1 c2_n$N; do python -c "
import json
d=json.load(open('$OUT/${TAG}_$f.json'))
print('$f', 'n_gpus', d['n_gpus'], 'value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], 'frac_of_ceiling', d['e2e'].get('frac_of_ceiling'), 'lat', (d.get('latency') or {}).get('single_configuration_us'))"; done
