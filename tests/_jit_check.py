"""Run in a subprocess with RFB200_JIT=2 (every eligible length goes through the NVRTC-specialised kernel)
and RFB200_JIT_VERBOSE=1: compares a set of smooth non-power-of-two transforms with the oracle
(the compiled reference when oracle/_ref is present, else the numpy restatement).  Exits non-zero on mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np

import parity
import rocket_fft_b200 as R
from oracle import pocketfft_oracle as O

T = parity.reflib() or O
rng = np.random.default_rng(11)
bad = 0
CD = {np.float32: np.complex64, np.float64: np.complex128}


def check(name, got, want, dtype, n):
    global bad
    e = parity.l2err(got, want)
    t = parity.tol(dtype, n)
    ok = e <= t
    bad += not ok
    print(f"{'ok ' if ok else 'BAD'} {name}: err {e:.2e} tol {t:.2e}", flush=True)


for dt in (np.complex64, np.complex128):
    for n in (6, 96, 100, 360, 1000, 1920, 3000, 6561, 15015):
        if dt == np.complex128 and n > 6561:
            continue
        rows = 5
        x = (rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n))).astype(dt)
        for fwd in (True, False):
            got, want = np.empty_like(x), np.empty_like(x)
            R.c2c(x, got, [1], fwd, 0.5)
            T.c2c(x, want, [1], fwd, 0.5, 1)
            check(f"c2c {np.dtype(dt).name} n={n} fwd={fwd}", got, want, dt, n)
    # strided lines (columns) and in-place
    x = (rng.standard_normal((360, 37)) + 1j * rng.standard_normal((360, 37))).astype(dt)
    want = np.empty_like(x)
    T.c2c(x, want, [0], True, 1.0, 1)
    got = x.copy()
    R.c2c(got, got, [0], True, 1.0)
    check(f"c2c columns in-place {np.dtype(dt).name} (360,37)", got, want, dt, 360)
# real transforms on jitted lengths
for dt, cdt in ((np.float32, np.complex64), (np.float64, np.complex128)):
    for n in (30, 1000, 1001, 1002, 1920):
        x = rng.standard_normal((7, n)).astype(dt)
        got = np.zeros((7, n // 2 + 1), dtype=cdt)
        want = np.zeros_like(got)
        for fwd in (True, False):  # even n: packed half-length transform + Hermitian unpack (kernel MODE 1)
            R.r2c(x, got, [1], fwd, 0.25)
            T.r2c(x, want, [1], fwd, 0.25, 1)
            check(f"r2c {np.dtype(dt).name} n={n} fwd={fwd}", got, want, dt, n)
        R.r2c(x, got, [1], True, 1.0)
        T.r2c(x, want, [1], True, 1.0, 1)
        back, wback = np.empty_like(x), np.empty_like(x)
        spec = want + (0.3 + 0.2j)  # non-zero imaginary parts in bin 0 / Nyquist must be ignored
        for fwd in (False, True):   # even n: Hermitian fold + half-length transform (kernel MODE 2)
            R.c2r(spec, back, [1], fwd, 1.0 / n)
            T.c2r(spec, wback, [1], fwd, 1.0 / n, 1)
            check(f"c2r {np.dtype(dt).name} n={n} fwd={fwd}", back, wback, dt, n)
        outs = np.zeros((7, 2 * n), dtype=dt)  # strided real output (every other slot)
        R.c2r(spec, outs[:, ::2], [1], False, 1.0)
        T.c2r(spec, wback, [1], False, 1.0, 1)
        check(f"c2r strided out {np.dtype(dt).name} n={n}", outs[:, ::2], wback, dt, n)
    x = rng.standard_normal((1080, 48)).astype(dt)
    got, want = np.empty_like(x), np.empty_like(x)
    R.r2r_separable_hartley(x, got, [0, 1], 1.0)
    T.r2r_separable_hartley(x, want, [0, 1], 1.0, 1)
    check(f"hartley {np.dtype(dt).name} (1080,48)", got, want, dt, 1080 * 48)
# zero-padded real lines on the device (n_in < n in the packed load) and a 2-D real transform
import torch

for dt in (np.float32, np.float64):
    x = rng.standard_normal((9, 700)).astype(dt)
    for n in (1000, 1080, 701):
        xp = np.zeros((9, n), dtype=dt)
        xp[:, :700] = x
        want = np.zeros((9, n // 2 + 1), dtype=CD[dt])
        T.r2c(xp, want, [1], True, 1.0, 1)
        got = torch.empty((9, n // 2 + 1), dtype=getattr(torch, np.dtype(CD[dt]).name), device="cuda")
        R.r2c_pad(torch.from_numpy(x).cuda(), got, (9, n), [1], True, 1.0)
        check(f"r2c_pad {np.dtype(dt).name} 700 -> {n}", got.cpu().numpy(), want, dt, n)
    x = rng.standard_normal((120, 360)).astype(dt)
    got = np.zeros((120, 181), dtype=CD[dt])
    want = np.zeros_like(got)
    R.r2c(x, got, [0, 1], True, 1.0)
    T.r2c(x, want, [0, 1], True, 1.0, 1)
    check(f"rfft2 {np.dtype(dt).name} (120,360)", got, want, dt, 120 * 360)
print("launches", R.launch_count())
sys.exit(1 if bad else 0)
