#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for cfg in "RFB200_FUSE4_BTMA=1" "RFB200_FUSE4_DEBUG_COPY=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=2" \
  "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=2 RFB200_FUSE4_BTMA=0" "RFB200_FUSE4_DEBUG_ONLY=1" "RFB200_FUSE4_DEBUG_ONLY=2" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=1 RFB200_FUSE4_CTAS=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=2 RFB200_FUSE4_CTAS=1"; do
  echo "-- $cfg"
  env RFB200_FUSE4_PF=0 RFB200_FUSE4_BTMA=1 $cfg timeout 120 python tools/microbench.py cfg2 2>&1 | grep "cols"
done 2>&1 | tee $O/r2g_sweep.log
RFB200_FUSE4_PF=0 RFB200_FUSE4_BTMA=1 RFB200_FUSE4_DEBUG_COPY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 1 -c 1 -o $O/r2g_copy python tools/prof_target.py cols 3 > $O/r2g_ncu.log 2>&1
