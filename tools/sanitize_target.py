"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rocket_fft_b200 as R

rng = np.random.default_rng(0)


def cplx(shape, dt):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)


for dt, cdt in ((np.float32, np.complex64), (np.float64, np.complex128)):
    # pow2 kernel: element-fast, line-fast, every pass structure; generic tile kernel; four-step; Bluestein
    for n in (1, 2, 3, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 30, 105, 1001, 15015, 67, 521, 32768, 65536, 70001):
        x = cplx((3, n), cdt)
        y = np.empty_like(x)
        R.c2c(x, y, [1], True, 1.0)
        assert np.allclose(y, np.fft.fft(x, axis=1), rtol=2e-3 if dt is np.float32 else 1e-9, atol=1e-2 if dt is np.float32 else 1e-8), n
        xt = np.ascontiguousarray(x.T)
        yt = np.empty_like(xt)
        R.c2c(xt, yt, [0], False, 1.0)
    x = cplx((20, 24, 18), cdt)
    y = np.empty_like(x)
    R.c2c(x[::-1, ::2, :], y[:, :12], [0, 1, 2], True, 1.0)
    # real transforms, fused r2c / c2r, Hartley, fftpack, c2c_sym
    for n in (1, 2, 31, 32, 64, 100, 4096, 16384, 4099):
        r = rng.standard_normal((5, n)).astype(dt)
        X = np.zeros((5, n // 2 + 1), dtype=cdt)
        R.r2c(r, X, [1], True, 1.0)
        back = np.empty_like(r)
        R.c2r(X, back, [1], False, 1.0 / n)
        assert np.allclose(back, r, atol=1e-3 if dt is np.float32 else 1e-9), n
        full = np.empty((5, n), dtype=cdt)
        R.c2c_sym(r, full, [1], True, 1.0)
        h = np.empty_like(r)
        R.r2r_separable_hartley(r, h, [1], 1.0)
        R.r2r_fftpack(r, h, [1], True, True, 1.0)
        R.r2r_fftpack(r, h, [1], False, False, 1.0)
    r = rng.standard_normal((12, 10, 9)).astype(dt)
    h = np.empty_like(r)
    R.r2r_genuine_hartley(r, h, [0, 1, 2], 1.0)
    X = np.zeros((12, 10, 5), dtype=cdt)
    R.r2c(r, X, [0, 1, 2], True, 1.0)
    R.c2r(X, h, [0, 1, 2], False, 1.0)
    # DCT / DST: fused power-of-two paths (contiguous + strided) and the generic path
    for shp, axes in (((4, 64), [1]), ((64, 8), [0]), ((4, 100), [1]), ((5, 2048), [1]), ((33, 3), [0])):
        r = rng.standard_normal(shp).astype(dt)
        for t in (1, 2, 3, 4):
            for kind in (R.dct, R.dst):
                o = np.empty_like(r)
                kind(r, o, axes, t, 1.0, True)
print("sanitize target done, launches:", R.launch_count())
