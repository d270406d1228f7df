#!/bin/bash
# Short 8-GPU visit (charged 8x): the default bench at 8 and 2 GPUs with the batch-256 default, and BASELINE config 4
# (C4, 64 walkers in total) sharded by walker over 8 GPUs after the rho split-heuristic fix.
N=${1:-8}
TAG=${2:-r02r}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | wc -l > $OUT/${TAG}_ngpu.txt
run() { # name, gpus, args...
    local name=$1 g=$2; shift 2
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g "$@" > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
    python -c "
import json,sys
d=json.load(open('$OUT/${TAG}_${name}.json'))
print('$name', 'n_gpus', d['n_gpus'], 'value %.0f' % d['value'], 'e2e %.0f' % (d['e2e']['value'] if d.get('e2e') else 0), 'frac_of_ceiling', (d['e2e'] or {}).get('frac_of_ceiling'), 'ceiling/gpu', (d['e2e'] or {}).get('h2d_ceiling_gbs'), d['scaling'])
" 2>&1 | tail -1
}
COMMON="--no-cpu-baseline --no-ab --no-pair --no-latency"
run c2_n$N $N --steps 20 --warmup 5 $COMMON
run c4_walker_n$N $N --workload C4 --total-batch 64 --steps 6 --warmup 3 $COMMON
run c2_n2 2 --steps 20 --warmup 5 $COMMON
