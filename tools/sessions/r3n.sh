#!/bin/bash
# Round 2, session 3n: N-D r2c through the padded scratch: parity, rfftn / irfftn 1024^3 with and without it.
set -u
O=gpurun_out
mkdir -p $O
( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "padded_scratch or golden or nd_layouts or r2c" ) > $O/r3n_pytest.log 2>&1
tail -5 $O/r3n_pytest.log
timeout -s KILL 300 python tools/microbench.py rfftn 2>&1 | tee $O/r3n_rfftn.log
