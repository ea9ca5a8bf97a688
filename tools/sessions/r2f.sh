#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
RFB200_FUSE4_AREG=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_fourstep" 2>&1 | tail -3 | tee $O/r2f_parity.log
for cfg in "RFB200_FUSE4_BTMA=1" "RFB200_FUSE4_AREG=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_BTMA=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_AREG=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_BTMA=1 RFB200_FUSE4_STAGES=3" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_BTMA=1 RFB200_FUSE4_PF=2" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_BTMA=1 RFB200_FUSE4_CTAS=2" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_BTMA=1 RFB200_FUSE4_CTAS=1"; do
  echo "-- $cfg"
  env RFB200_FUSE4_PF=0 $cfg timeout 120 python tools/microbench.py cfg2 2>&1 | grep "cols"
done 2>&1 | tee $O/r2f_sweep.log
