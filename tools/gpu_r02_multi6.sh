#!/bin/bash
# 8-GPU check of the new default (bins of 16 x 512): the driver's command, short.
OUT=gpurun_out
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-ab --no-pair > $OUT/r02last_c2_n8.json 2> $OUT/r02last_c2_n8.err
python -c "
import json
d=json.load(open('$OUT/r02last_c2_n8.json'))
print('n8', d['n_gpus'], 'value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], d['e2e'].get('frac_of_ceiling'), (d.get('latency') or {}).get('single_configuration_us'))" || tail -5 $OUT/r02last_c2_n8.err
