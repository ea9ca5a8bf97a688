#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py tests/test_gpu_fft_api.py tests/test_gpu_distributed.py -x -q -m gpu --durations=8 -k "not full_size_against" ) > $O/r2j_pytest_gpu.log 2>&1
tail -25 $O/r2j_pytest_gpu.log
timeout 300 python tools/microbench.py cfg5 2>&1 | tee $O/r2j_cfg5.log
timeout 200 python - <<'PY' 2>&1 | tee $O/r2j_dct_types.log
import torch, rocket_fft_b200 as R
from tools.microbench import timeit
x = torch.randn(2048, 2048, 64, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
nb = 2 * x.numel() * 8
for kind in ("dct", "dst"):
    for typ in (1, 2, 3, 4):
        ms = timeit(lambda: getattr(R, kind)(x, y, [0, 1], typ, 1.0, False), 3)
        print(f"{kind} type {typ} f64 (2048,2048,64) axes (0,1): {ms:.3f} ms  {nb/ms/1e6:.0f} GB/s")
x = torch.randn(4096, 4097, dtype=torch.float32, device="cuda"); y = torch.empty_like(x)
for typ in (1, 2, 3, 4):
    ms = timeit(lambda: R.dct(x, y, [1], typ, 1.0, False), 3)
    print(f"dct type {typ} f32 (4096,4097) axis 1: {ms:.3f} ms  {2*x.numel()*4/ms/1e6:.0f} GB/s")
PY
