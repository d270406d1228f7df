#!/bin/bash
# 8-GPU confirmation of the pipelined bin exchange: comm tests, C2 at 1 and N GPUs, C4 (64 walkers in total) at N.
N=${1:-8}
TAG=${2:-r02t}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_comm.py -m gpu -x -q 2>&1 | tail -4 > $OUT/${TAG}_pytest_comm.log; cat $OUT/${TAG}_pytest_comm.log
run() { # name, gpus, args...
    local name=$1 g=$2; shift 2
    if [ "$g" = "1" ]; then
        timeout 600 python bench.py --gpus 1 "$@" > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
    else
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g "$@" > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
    fi
    python -c "
import json,sys
d=json.load(open('$OUT/${TAG}_${name}.json'))
print('$name', 'n_gpus', d['n_gpus'], 'value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % (d['e2e']['value'] if d.get('e2e') else 0), 'frac_of_ceiling', (d['e2e'] or {}).get('frac_of_ceiling'), d['scaling'])
" 2>&1 | tail -1
}
COMMON="--no-cpu-baseline --no-ab --no-pair --no-latency"
run c2_n1 1 --steps 40 --warmup 5 $COMMON
run c2_n${N} $N --steps 40 --warmup 5 $COMMON
run c4_walker_n${N} $N --workload C4 --total-batch 64 --steps 6 --warmup 3 $COMMON
