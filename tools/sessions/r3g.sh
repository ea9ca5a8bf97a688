#!/bin/bash
# Round 2, session 3g: two-lines-per-thread kernel with 16-byte pair accesses: parity and cfg3 per axis.
set -u
O=gpurun_out
mkdir -p $O
( timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_per_thread or streamed or slab or nd_layouts or strided" ) > $O/r3g_pytest.log 2>&1
tail -4 $O/r3g_pytest.log
timeout -s KILL 200 python tools/microbench.py cfg3 2>&1 | tee $O/r3g_cfg3.log
