#!/bin/bash
# Round 2, session x: fused column kernel with sector-aligned B windows: parity, timing, copy modes.
set -u
O=gpurun_out
mkdir -p $O
( timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_fourstep or full_size" ) > $O/r2x_pytest.log 2>&1
tail -15 $O/r2x_pytest.log
for env in "RFB200_FUSE4=1" "RFB200_FUSE4_DEBUG_COPY=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=1" "RFB200_FUSE4_DEBUG_COPY=1 RFB200_FUSE4_DEBUG_ONLY=2"; do
  echo "-- $env"
  env $env timeout -s KILL 200 python tools/probe_fused_rows.py 2>&1 | grep -v "^rocketfft"
done | tee $O/r2x_fused_aligned_windows.log
timeout -s KILL 200 python tools/microbench.py cfg2 2>&1 | tee $O/r2x_cfg2.log
