#!/bin/bash
# GPU session: c2c mode of the two-transforms-per-thread kernel (8192 / 16384-point complex64 lines): parity and A/B.
set -u
O=gpurun_out
mkdir -p $O
echo "== parity, RFB200_DUAL_C2C=1 (both lengths on)"
RFB200_DUAL_C2C=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_complex or selected_lengths or every_length or golden" 2>&1 | tail -8 | tee $O/r1h_pytest_dualc2c.log
echo "== parity, defaults (8192 on, 16384 off)"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_complex or selected_lengths" 2>&1 | tail -8 | tee -a $O/r1h_pytest_dualc2c.log
cat > /tmp/ab.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, rocket_fft_b200 as R
dev = torch.device("cuda:0")
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for n in (4096, 8192, 16384):
    rows = (1 << 28) // (n * 8)
    x = torch.randn(rows, n, dtype=torch.complex64, device=dev); y = torch.empty_like(x)
    ms = timeit(lambda: R.c2c(x, y, [1], True, 1.0))
    cu = timeit(lambda: torch.fft.fft(x, dim=1))
    gb = 2 * x.numel() * 8 / ms / 1e6
    print(f"DUAL_C2C={os.environ.get('RFB200_DUAL_C2C','default')} c2c c64 ({rows},{n}): {ms:.4f} ms {gb:.0f} GB/s {gb/6527.8*100:.1f}% of measured peak | cuFFT {cu:.4f} ms", flush=True)
P
for v in 0 1; do RFB200_DUAL_C2C=$v timeout 120 python /tmp/ab.py; done 2>&1 | tee $O/r1h_ab_dual_c2c.log
echo "== done"
