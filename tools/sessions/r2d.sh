#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_fourstep" 2>&1 | tail -5 | tee $O/r2d_parity.log
for cfg in "RFB200_FUSE4=2 RFB200_FUSE4_PF=0" "RFB200_FUSE4=2 RFB200_FUSE4_PF=2" "RFB200_FUSE4=2 RFB200_FUSE4_PF=0 RFB200_FUSE4_STAGES=3" "RFB200_FUSE4=2 RFB200_FUSE4_PF=0 RFB200_FUSE4_RING=14 RFB200_FUSE4_LAG=8"; do
  echo "-- $cfg"
  env $cfg timeout 120 python tools/microbench.py cfg2 2>&1 | grep -v "cuFFT\|^rocketfft"
done 2>&1 | tee $O/r2d_sweep.log
