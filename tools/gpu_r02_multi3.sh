#!/bin/bash
# N-GPU visit for the pipelined bin exchange: comm tests, then the default bench with both exchange modes.
N=${1:-2}
TAG=${2:-r02s}
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
    tail -3 $OUT/${TAG}_${name}.err | cut -c1-300
}
COMMON="--no-cpu-baseline --no-ab --no-pair --no-latency"
run c2_n1 1 --steps 40 --warmup 5 $COMMON
run c2_n${N}_pipelined $N --steps 40 --warmup 5 $COMMON
run c2_n${N}_inline $N --steps 40 --warmup 5 --exchange inline $COMMON
