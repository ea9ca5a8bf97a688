#!/bin/bash
# Round 2, session l: robustness tests, L2-blocked axis groups (cfg3, batched 2-D), host copy probe.
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_distributed.py -x -q -m gpu ) > $O/r2l_pytest.log 2>&1
tail -5 $O/r2l_pytest.log
for mb in 0 16 32 64; do
  echo "-- RFB200_L2BLOCK_MB=$mb"
  RFB200_L2BLOCK_MB=$mb timeout 300 python tools/microbench.py cfg3 batch2d 2>&1
done | tee $O/r2l_l2block.log
timeout 300 python tools/host_copy_probe.py 2>&1 | tee $O/r2l_host_copy.log
