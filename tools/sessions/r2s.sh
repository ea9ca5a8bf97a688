#!/bin/bash
# Round 2, session s: fused column kernel vs the distance between a tile's rows (page spread), copy modes.
set -u
O=gpurun_out
mkdir -p $O
for env in "RFB200_FUSE4=1" "RFB200_FUSE4_DEBUG_COPY=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=2" "RFB200_FUSE4=0"; do
  echo "-- $env"
  env $env timeout -s KILL 200 python tools/probe_fused_rows.py 2>&1 | grep -v "^rocketfft"
done | tee $O/r2s_fused_rows.log
