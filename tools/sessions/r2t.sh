#!/bin/bash
# Round 2, session t: fused column kernel, L2 policy of the array stores (evict_first / normal / last); B-only copy mode.
set -u
O=gpurun_out
mkdir -p $O
for pol in 0 1 2; do
 for env in "RFB200_FUSE4=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=2"; do
  echo "-- RFB200_FUSE4_STPOL=$pol $env"
  env RFB200_FUSE4_STPOL=$pol $env timeout -s KILL 200 python tools/probe_fused_rows.py 2>&1 | grep -v "^rocketfft" | head -2
 done
done | tee $O/r2t_fused_stpol.log
