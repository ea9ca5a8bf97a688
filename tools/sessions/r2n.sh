#!/bin/bash
# Round 2, session n: the streamed strided-lines kernel (parity, cfg3), the L2-blocked four-step experiment, staging sweep.
set -u
O=gpurun_out
mkdir -p $O
( timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed or two_per_thread" ) > $O/r2n_pytest.log 2>&1
tail -15 $O/r2n_pytest.log
for env in "RFB200_STREAM=0" "RFB200_STREAM=1" "RFB200_STREAM=0 RFB200_LF4=1" "RFB200_STREAM=0 RFB200_LF4=1 RFB200_LF4_MB=64"; do
  echo "-- $env"
  env $env RFB200_L2BLOCK_MB=0 timeout -s KILL 200 python tools/microbench.py cfg3 2>&1
done | tee $O/r2n_cfg3.log
for env in "RFB200_STAGE_THREADS=8" "RFB200_STAGE_THREADS=12" "RFB200_STAGE_THREADS=12 RFB200_STAGE_SLOT_MB=16" "RFB200_STAGE_THREADS=14 RFB200_STAGE_SLOT_MB=64"; do
  echo "-- $env"
  env $env timeout -s KILL 200 python tools/stage_probe.py 2 2>&1
done | tee $O/r2n_stage.log
