#!/bin/bash
# Multi-GPU record visit: two-rank tests of the library's own NCCL entry points, the default bench at N GPUs (walker
# sharding, library collective), and BASELINE config 4 (C4) sharded by q and by walker.
N=${1:-8}
TAG=${2:-r02m}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | wc -l > $OUT/${TAG}_ngpu.txt
timeout 600 python -m pytest tests/test_comm.py tests/test_host_layer.py -m gpu -x -q -k "comm or two_ranks or batched_device_bins or rank" 2>&1 | tail -6 > $OUT/${TAG}_pytest_multi.log; cat $OUT/${TAG}_pytest_multi.log
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
print('$name', 'n_gpus', d['n_gpus'], 'value %.0f' % d['value'], 'e2e %.0f' % (d['e2e']['value'] if d.get('e2e') else 0), 'frac_of_ceiling', (d['e2e'] or {}).get('frac_of_ceiling'), 'ceiling/gpu', (d['e2e'] or {}).get('h2d_ceiling_gbs'), d['scaling'])
" 2>&1 | tail -1
}
COMMON="--no-cpu-baseline --no-ab --no-pair --no-latency"
run c2_n1 1 --steps 20 --warmup 5 $COMMON
for g in 2 4 8; do [ $g -le $N ] && run c2_n$g $g --steps 20 --warmup 5 $COMMON; done
run c4_n1 1 --workload C4 --total-batch 64 --steps 6 --warmup 3 $COMMON
run c4_walker_n$N $N --workload C4 --total-batch 64 --steps 6 --warmup 3 $COMMON
run c4_qshard_n$N $N --workload C4 --shard q --batch 64 --steps 6 --warmup 3 $COMMON
run c4_qshard_b8_n$N $N --workload C4 --shard q --batch 8 --steps 6 --warmup 3 $COMMON
run c4_b8_n1 1 --workload C4 --batch 8 --steps 6 --warmup 3 $COMMON
