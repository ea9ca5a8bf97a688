#!/bin/bash
# Round 2, session c: fused four-step kernel with tensor-map TMA: parity, timing sweeps, one ncu capture.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_fourstep or strided_lines" 2>&1 | tail -15 | tee $O/r2c_parity.log
for cfg in "RFB200_FUSE4=0" "RFB200_FUSE4=2" "RFB200_FUSE4=2 RFB200_FUSE4_PF=0" "RFB200_FUSE4=2 RFB200_FUSE4_PF=4" "RFB200_FUSE4=2 RFB200_FUSE4_STAGES=3" "RFB200_FUSE4=2 RFB200_FUSE4_RING=14 RFB200_FUSE4_LAG=8" \
           "RFB200_FUSE4=2 RFB200_FUSE4_RING=6 RFB200_FUSE4_LAG=3" "RFB200_FUSE4=2 RFB200_FUSE4_STAGES=3 RFB200_FUSE4_RING=14 RFB200_FUSE4_LAG=8"; do
  echo "-- $cfg"
  env $cfg timeout 120 python tools/microbench.py cfg2 2>&1 | grep -v "cuFFT\|^rocketfft"
done 2>&1 | tee $O/r2c_sweep.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused2 -s 1 -c 1 -o $O/r2c_cols_fused2 python tools/prof_target.py cols 3 > $O/r2c_ncu.log 2>&1
tail -3 $O/r2c_ncu.log
