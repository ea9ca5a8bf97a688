#!/bin/bash
# Round 2, session 3p: two-launch four-step with padded scratch rows: parity and timing (RFB200_FUSE4=0).
set -u
O=gpurun_out
mkdir -p $O
( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_fourstep or two_per_thread or selected_lengths or config4 or bluestein" ) > $O/r3p_pytest.log 2>&1
tail -4 $O/r3p_pytest.log
RFB200_FUSE4=0 timeout -s KILL 200 python tools/microbench.py cfg2 2>&1 | grep -E "cols|whole|irfft" | tee $O/r3p_cfg2_twolaunch.log
timeout -s KILL 200 python tools/microbench.py cfg4 2>&1 | tee -a $O/r3p_cfg2_twolaunch.log
