#!/bin/bash
# Mini session: row-wise L2 prefetch for the second four-step pass of the rfft2 columns (rows 65544 B apart, just above the
# 64 KiB cut-off of set_prefetch_rows).
set -u
O=gpurun_out
mkdir -p $O
cat > /tmp/cols.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, rocket_fft_b200 as R
dev = torch.device("cuda:0")
X = torch.randn(16384, 8193, dtype=torch.complex64, device=dev)
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = timeit(lambda: R.c2c(X, X, [0], True, 1.0))
print(f"MAXSTRIDE={os.environ.get('RFB200_PF_LF_MAXSTRIDE','default')} PF_LF={os.environ.get('RFB200_PF_LF','default')} cols 16384x8193: {ms:.4f} ms", flush=True)
P
( timeout 40 python /tmp/cols.py
  RFB200_PF_LF_MAXSTRIDE=70000 timeout 40 python /tmp/cols.py
  RFB200_PF_LF_MAXSTRIDE=70000 RFB200_PF_LF=8 timeout 40 python /tmp/cols.py
  RFB200_PF_LF_MAXSTRIDE=70000 RFB200_PF_LF=96 timeout 40 python /tmp/cols.py ) 2>&1 | tee $O/r1j_cols_prefetch_rows.log
