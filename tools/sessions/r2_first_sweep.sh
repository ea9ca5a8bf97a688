#!/bin/bash
# Prepared for the next round (nothing here has been run yet): the experiments left open at the end of round 1.
#   1. pure-copy ceilings of the strided passes (tools/probe_strided_copy.py)
#   2. 128-point strided kernel compiled for 5 / 6 resident CTAs per SM (RFB200_LF_CTAS), alone and under the fused
#      four-step kernel (RFB200_FUSE4=1) -- parity first, then timings
# Usage: gpurun --timeout 600 -- 'bash tools/sessions/r2_first_sweep.sh'
set -u
O=gpurun_out
mkdir -p $O
timeout 60 python tools/probe_strided_copy.py 2>&1 | tee $O/r2a_probe.log
for ctas in 5 6; do
  echo "== parity RFB200_LF_CTAS=$ctas"
  RFB200_LF_CTAS=$ctas timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "strided_lines or fused_fourstep or device_arrays or golden" 2>&1 | tail -4 | tee -a $O/r2a_parity_lf_ctas.log
done
for cfg in "RFB200_LF_CTAS=4" "RFB200_LF_CTAS=5" "RFB200_LF_CTAS=6" "RFB200_FUSE4=1 RFB200_FUSE4_COLS=128" "RFB200_FUSE4=1 RFB200_FUSE4_COLS=128 RFB200_LF_CTAS=5" \
           "RFB200_FUSE4=1 RFB200_FUSE4_COLS=128 RFB200_LF_CTAS=6" "RFB200_FUSE4=1 RFB200_FUSE4_COLS=64 RFB200_FUSE4_RING=6 RFB200_FUSE4_LAG=3 RFB200_LF_CTAS=6"; do
  echo "-- $cfg"
  env $cfg timeout 120 python tools/microbench.py cfg2 2>&1 | grep -v "cuFFT\|^rocketfft"
done 2>&1 | tee $O/r2a_sweep_lf_ctas.log
