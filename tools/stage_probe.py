"""Measurement aid: rfft2 of float32 16384^2 images through numba_r2c on pageable NumPy arrays vs pinned buffers
(the staging ring of csrc/staging.cu; RFB200_STAGE_THREADS / RFB200_STAGE_SLOT_MB).  Usage: python tools/stage_probe.py [images]"""
import sys
import time

import numpy as np
import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rocket_fft_b200 as R

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H = W = 16384
px = np.empty((B, H, W), dtype=np.float32)
px[...] = 0.5
pX = np.empty((B, H, W // 2 + 1), dtype=np.complex64)
pX[...] = 0


def wall(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


dt = wall(lambda: R.r2c(px, pX, [1, 2], True, 1.0))
print(f"pageable: {dt * 1e3:8.1f} ms per step of {B} images  ({(px.nbytes + pX.nbytes) / dt / 1e9:.1f} GB/s both directions)", flush=True)
hx = torch.empty(B, H, W, dtype=torch.float32, pin_memory=True)
hX = torch.empty(B, H, W // 2 + 1, dtype=torch.complex64, pin_memory=True)
nx, nX = hx.numpy(), hX.numpy()
nx[...] = 0.5
dt2 = wall(lambda: R.r2c(nx, nX, [1, 2], True, 1.0))
print(f"pinned:   {dt2 * 1e3:8.1f} ms  -> pageable / pinned = {dt2 / dt:.3f}", flush=True)
