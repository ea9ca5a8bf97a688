#!/bin/bash
# Round 2, session 3o: rfftn / irfftn 1024^3 and irfft2 16384^2 with and without the padded scratch rows (A/B).
set -u
O=gpurun_out
mkdir -p $O
for v in 1 0; do
  echo "-- RFB200_NO_PADDED_SCRATCH=$v"
  RFB200_NO_PADDED_SCRATCH=$v timeout -s KILL 300 python tools/microbench.py rfftn cfg2 2>&1 | grep -E "rfftn|irfft"
done | tee $O/r3o_padded_ab.log
