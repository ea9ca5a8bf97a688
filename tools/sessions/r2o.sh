#!/bin/bash
# Round 2, session o: streamed strided-lines kernel, cp.async variant (parity, cfg3 per axis), staging sweep.
set -u
O=gpurun_out
mkdir -p $O
( RFB200_STREAM_CP=1 timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed" ) > $O/r2o_pytest_cp.log 2>&1
tail -5 $O/r2o_pytest_cp.log
( timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed or two_per_thread or slab" ) > $O/r2o_pytest.log 2>&1
tail -5 $O/r2o_pytest.log
for env in "RFB200_STREAM_CP=1" "RFB200_STREAM_CP=-1" ; do
  echo "-- $env"
  env $env timeout -s KILL 200 python tools/microbench.py cfg3 2>&1
done | tee $O/r2o_cfg3.log
for env in "RFB200_STAGE_THREADS=8" "RFB200_STAGE_THREADS=12" "RFB200_STAGE_THREADS=12 RFB200_STAGE_SLOT_MB=16" "RFB200_STAGE_THREADS=14 RFB200_STAGE_SLOT_MB=64"; do
  echo "-- $env"
  env $env timeout -s KILL 200 python tools/stage_probe.py 2 2>&1
done | tee $O/r2o_stage.log
