#!/bin/bash
# Two-GPU visit: the library's own NCCL exchange step (tests + bench A/B against torch.distributed on the same buffers).
TAG=${1:-g2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | tee $OUT/${TAG}_gpu.txt
python -m pytest tests/test_comm.py -m gpu -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_comm.log
BA="--no-cpu-baseline --no-ab --no-pair --no-latency --steps 200 --warmup 10"
for coll in torch lib; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $BA --collective $coll > $OUT/${TAG}_bench_${coll}.json 2> $OUT/${TAG}_bench_${coll}.err
  tail -c 1500 $OUT/${TAG}_bench_${coll}.json; tail -2 $OUT/${TAG}_bench_${coll}.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $BA --no-e2e --collective lib --shard q > $OUT/${TAG}_bench_lib_q.json 2> $OUT/${TAG}_bench_lib_q.err
tail -c 600 $OUT/${TAG}_bench_lib_q.json; tail -2 $OUT/${TAG}_bench_lib_q.err
