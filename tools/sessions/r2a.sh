#!/bin/bash
# Round 2, session a: probes that decide the design of the fused column transform + the experiments left open in round 1.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/r2a_smi.txt 2>&1
timeout 200 python tools/probe_l2_dsmem.py 2>&1 | tee $O/r2a_probe_l2_dsmem.log
timeout 100 python tools/probe_strided_copy.py 2>&1 | tee $O/r2a_probe_strided.log
for ctas in 5 6; do
  echo "== parity RFB200_LF_CTAS=$ctas"
  RFB200_LF_CTAS=$ctas timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "strided_lines or fused_fourstep or golden" 2>&1 | tail -3
done 2>&1 | tee $O/r2a_parity_lf_ctas.log
for cfg in "RFB200_LF_CTAS=4" "RFB200_LF_CTAS=5" "RFB200_LF_CTAS=6" "RFB200_FUSE4=1 RFB200_FUSE4_COLS=128" "RFB200_FUSE4=1 RFB200_FUSE4_COLS=128 RFB200_LF_CTAS=5" \
           "RFB200_FUSE4=1 RFB200_FUSE4_COLS=128 RFB200_LF_CTAS=6"; do
  echo "-- $cfg"
  env $cfg timeout 120 python tools/microbench.py cfg2 2>&1 | grep -v "cuFFT\|^rocketfft"
done 2>&1 | tee $O/r2a_sweep_lf_ctas.log
