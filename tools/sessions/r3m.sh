#!/bin/bash
# Round 2, session 3m: c2r with a padded temporary: parity (c2r tests, full-size configs), irfft2 timing, single-rank slab check.
set -u
O=gpurun_out
mkdir -p $O
( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_distributed.py tests/test_gpu_fft_api.py -x -q -m gpu -k "r2c or c2r or real or full_size or golden or slab or irfft or fft_api or config2" ) > $O/r3m_pytest.log 2>&1
tail -5 $O/r3m_pytest.log
timeout -s KILL 200 python tools/microbench.py cfg2 2>&1 | tee $O/r3m_cfg2.log
