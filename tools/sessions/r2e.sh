#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for m in "RFB200_FUSE4_BTMA=0" "RFB200_FUSE4_AREG=1"; do
env $m timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_fourstep" 2>&1 | tail -3
done | tee $O/r2e_parity.log
for cfg in "RFB200_FUSE4_BTMA=1" "RFB200_FUSE4_BTMA=0" "RFB200_FUSE4_AREG=1" "RFB200_FUSE4_AREG=1 RFB200_FUSE4_STAGES=3" "RFB200_FUSE4_BTMA=0 RFB200_FUSE4_STAGES=3" "RFB200_FUSE4_AREG=1 RFB200_FUSE4_PF=2" "RFB200_FUSE4_BTMA=0 RFB200_FUSE4_PF=2"; do
  echo "-- $cfg"
  env RFB200_FUSE4_PF=0 $cfg timeout 120 python tools/microbench.py cfg2 2>&1 | grep "cols"
done 2>&1 | tee $O/r2e_sweep.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 1 -c 1 -o $O/r2e_cols_fused2 python tools/prof_target.py cols 3 > $O/r2e_ncu.log 2>&1
tail -2 $O/r2e_ncu.log
