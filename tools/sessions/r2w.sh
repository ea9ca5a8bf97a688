#!/bin/bash
# Round 2, session w: fused column kernel on arrays whose rows are 32-byte aligned (pitch 8196 / 8192 elements) vs 8193.
set -u
O=gpurun_out
mkdir -p $O
for env in "RFB200_FUSE4=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=2" "RFB200_FUSE4=0"; do
  echo "-- $env"
  env $env timeout -s KILL 200 python tools/probe_fused_rows.py 2>&1 | grep -v "^rocketfft"
done | tee $O/r2w_fused_pitch.log
