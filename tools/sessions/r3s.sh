#!/bin/bash
# Round 2, session 3s (8 GPUs): the driver's scaling command at N=8 (bench.py with the slab fftn inside the line).
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 ) > $O/r3s_bench_8gpu.json 2> $O/r3s_bench_8gpu.err
tail -c 3000 $O/r3s_bench_8gpu.json; tail -5 $O/r3s_bench_8gpu.err
