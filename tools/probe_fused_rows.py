"""Measurement aid for the fused four-step column kernel: does the distance between the rows of a tile matter?  The same number of
16384-point lines once as ONE wide array (rows 65544 B apart: the 128 rows of a tile lie 8.4 MB apart, a page each) and once
as a batch of narrow arrays (rows 2112 B apart: a tile's rows lie 270 KB apart, 8 rows per 2 MiB page).  Run with
RFB200_FUSE4_DEBUG_COPY / RFB200_FUSE4_DEBUG_ONLY (read once per process) for the A-only / B-only copy modes.
Usage: python tools/probe_fused_rows.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rocket_fft_b200 as R
from tools.microbench import timeit

dev = torch.device("cuda:0")
for shape, axis in (((16384, 8193), 0), ((31, 16384, 264), 1), ((8, 16384, 1025), 1), ((16384, 8193, 8196), 0), ((16384, 8192, 8192), 0)):
    if len(shape) == 3 and axis == 0:  # (rows, columns used, row pitch in elements): a view with 32-byte aligned rows
        x = torch.randn(shape[0], shape[2], dtype=torch.complex64, device=dev)[:, :shape[1]]
    else:
        x = torch.randn(*shape, dtype=torch.complex64, device=dev)
    R.launch_trace(True)
    R.c2c(x, x, [axis], True, 1.0)
    torch.cuda.synchronize()
    names = R.launch_trace_get()
    R.launch_trace(False)
    ms = timeit(lambda: R.c2c(x, x, [axis], True, 1.0), 5)
    nb = 2 * x.numel() * 8
    print(f"{str(shape):22s} axis {axis}: {ms:.4f} ms  {nb / ms / 1e6:7.0f} GB/s (algorithmic)  kernels={names}", flush=True)
    del x
