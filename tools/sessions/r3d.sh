#!/bin/bash
# Round 2, session 3d: single-rank slab check (runs on a 1-GPU box), randomised parity fuzz against the compiled reference.
set -u
O=gpurun_out
mkdir -p $O
( time timeout -s KILL 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu ) > $O/r3d_pytest_dist.log 2>&1
tail -12 $O/r3d_pytest_dist.log
( timeout -s KILL 400 python tools/fuzz_parity.py 150 11 ) > $O/r3d_fuzz.log 2>&1
tail -5 $O/r3d_fuzz.log
