#!/bin/bash
# Round 2, session p (2 GPUs): bench.py at N=2 with the slab fftn inside the line; 2-GPU slab tests.
set -u
O=gpurun_out
mkdir -p $O
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e ) > $O/r2p_bench_2gpu.json 2> $O/r2p_bench_2gpu.err
tail -c 1800 $O/r2p_bench_2gpu.json; tail -5 $O/r2p_bench_2gpu.err
( time timeout -s KILL 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu ) > $O/r2p_pytest_2gpu.log 2>&1
tail -5 $O/r2p_pytest_2gpu.log
