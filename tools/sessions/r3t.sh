#!/bin/bash
# Round 2, session 3t (4 GPUs): the driver's scaling command at N=4.
set -u
O=gpurun_out
mkdir -p $O
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 ) > $O/r3t_bench_4gpu.json 2> $O/r3t_bench_4gpu.err
tail -c 600 $O/r3t_bench_4gpu.json; tail -3 $O/r3t_bench_4gpu.err
