"""Tiny driver for ncu captures: runs one named workload a few times on cuda:0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rocket_fft_b200 as R

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if name == "cfg1":
    x = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
    y = torch.empty_like(x)
    fn = lambda: R.c2c(x, y, [1], True, 1.0)
elif name == "cfg1f":
    x = torch.randn(8192, 4096, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    fn = lambda: R.c2c(x, y, [1], True, 1.0)
elif name == "rows":
    x = torch.randn(16384, 16384, dtype=torch.float32, device=dev)
    y = torch.empty(16384, 8193, dtype=torch.complex64, device=dev)
    fn = lambda: R.r2c(x, y, [1], True, 1.0)
elif name == "cols":
    y = torch.randn(16384, 8193, dtype=torch.complex64, device=dev)
    fn = lambda: R.c2c(y, y, [0], True, 1.0)
elif name == "rfft2":
    x = torch.randn(16384, 16384, dtype=torch.float32, device=dev)
    y = torch.empty(16384, 8193, dtype=torch.complex64, device=dev)
    fn = lambda: R.r2c(x, y, [0, 1], True, 1.0)
elif name == "cfg3":
    x = torch.randn(512, 1024, 1024, dtype=torch.complex64, device=dev)
    fn = lambda: R.c2c(x, x, [0, 1, 2], True, 1.0)
elif name == "cfg3_axis0":
    x = torch.randn(1024, 512, 1024, dtype=torch.complex64, device=dev)  # 1024-point lines, rows 4 MiB apart
    fn = lambda: R.c2c(x, x, [0], True, 1.0)
elif name == "cfg4a":
    x = torch.randn(4096, 15015, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    fn = lambda: R.c2c(x, y, [1], True, 1.0)
elif name == "cfg4b":
    x = torch.randn(256, 1000003, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    fn = lambda: R.c2c(x, y, [1], True, 1.0)
elif name == "cfg5":
    x = torch.randn(2048, 2048, 64, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    fn = lambda: R.dct(x, y, [0, 1], 2, 1.0, False)
else:
    raise SystemExit("unknown workload")
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("done", name)
