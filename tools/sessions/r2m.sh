#!/bin/bash
# Round 2, session m (2 GPUs): slab fftn / rfftn against the oracle with all exchange engines, bench.py at N=2.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L | tee $O/r2m_smi.txt
( time timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu ) > $O/r2m_pytest_2gpu.log 2>&1
tail -15 $O/r2m_pytest_2gpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > $O/r2m_bench_2gpu.json 2> $O/r2m_bench_2gpu.err
tail -c 2500 $O/r2m_bench_2gpu.json; tail -5 $O/r2m_bench_2gpu.err
