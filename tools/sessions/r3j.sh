#!/bin/bash
# Round 2, session 3j: ring / lag sweep of the shipped fused column kernel.
set -u
O=gpurun_out
mkdir -p $O
for env in "RFB200_FUSE4_RING=10 RFB200_FUSE4_LAG=6" "RFB200_FUSE4_RING=12 RFB200_FUSE4_LAG=7" "RFB200_FUSE4_RING=12 RFB200_FUSE4_LAG=8" "RFB200_FUSE4_RING=14 RFB200_FUSE4_LAG=8" "RFB200_FUSE4_RING=14 RFB200_FUSE4_LAG=10" "RFB200_FUSE4_RING=9 RFB200_FUSE4_LAG=6" "RFB200_FUSE4_RING=8 RFB200_FUSE4_LAG=5" "RFB200_FUSE4_RING=11 RFB200_FUSE4_LAG=7" "RFB200_FUSE4_RING=16 RFB200_FUSE4_LAG=10"; do
  echo "-- $env"
  env $env timeout -s KILL 200 python tools/probe_fused_rows.py 2>&1 | grep -E "8193\)"
  env $env timeout -s KILL 200 python tools/probe_fused_rows.py 2>&1 | grep -E "8193\)"
done | tee $O/r3j_ring_lag.log
