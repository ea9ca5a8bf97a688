"""Host-side argument handling of the numpy.fft / scipy.fft style layer (rocket_fft_b200/fft.py): shape / axes
normalisation, norm -> fct, frequency helpers, next_fast_len.  No GPU needed: nothing here launches a kernel
(the reference's counterparts: rocket_fft/overloads.py:421-489, 505-551, 1862-1870)."""
import math

import numpy as np
import pytest
import scipy.fft

from rocket_fft_b200 import fft as F


def test_shape_and_axes_normalisation():
    x = np.zeros((4, 5, 6))
    assert F._shape_axes(x, None, None, True) == ([4, 5, 6], [0, 1, 2])
    assert F._shape_axes(x, None, None, False) == ([6], [2])
    assert F._shape_axes(x, (8, 3), None, True) == ([8, 3], [1, 2])
    assert F._shape_axes(x, None, (-1, 0), True) == ([6, 4], [2, 0])
    assert F._shape_axes(x, (7, -1), (0, 1), True) == ([7, 5], [0, 1])
    assert F._shape_axes(x, 9, 1, True) == ([9], [1])
    with pytest.raises(ValueError):
        F._shape_axes(x, None, (3,), True)
    with pytest.raises(ValueError):
        F._shape_axes(x, (1, 2, 3), (0, 1), True)
    with pytest.raises(ValueError):
        F._shape_axes(x, (0,), (0,), True)
    assert F._target_shape(x, [8, 3], [1, 2]) == [4, 8, 3]
    assert F._needs_pad(x, [4, 8, 3]) and not F._needs_pad(x, [4, 5, 3])
    # cropping is a view, zero-padding a copy
    y = np.arange(24.0).reshape(4, 6)
    c = F._pad_or_crop(y, [3], [1], np.float64)
    assert c.shape == (4, 3) and np.shares_memory(c, y)
    p = F._pad_or_crop(y, [8, 2], [1, 0], np.float64)
    assert p.shape == (2, 8) and not np.shares_memory(p, y) and np.array_equal(p[:, :6], y[:2]) and not p[:, 6:].any()


def test_norm_factors():
    shape = (4, 5, 6)
    for axes in ([2], [0, 1], [0, 1, 2], [1, 1]):
        n = math.prod(shape[a] for a in axes)
        assert F._fct(shape, axes, None, True) == 1.0 and F._fct(shape, axes, "backward", False) == 1.0 / n
        assert F._fct(shape, axes, "forward", True) == 1.0 / n and F._fct(shape, axes, "forward", False) == 1.0
        assert F._fct(shape, axes, "ortho", True) == F._fct(shape, axes, "ortho", False) == 1.0 / math.sqrt(n)
    # DCT/DST: logical length 2(N + delta) per axis (delta = -1 DCT-I, +1 DST-I, 0 otherwise)
    assert F._fct(shape, [1], None, False, -1.0) == 1.0 / (2.0 * 4)
    assert F._fct(shape, [1], None, False, 1.0) == 1.0 / (2.0 * 6)
    assert F._fct(shape, [0, 2], "ortho", True, 0.0) == 1.0 / math.sqrt(8.0 * 12.0)
    with pytest.raises(ValueError):
        F._fct(shape, [0], "nope", True)


def test_dtype_promotion_rules():
    assert F._real_of(np.float16) == np.float32 and F._real_of(np.complex64) == np.float32
    assert F._real_of(np.int64) == np.float64 and F._real_of(np.complex128) == np.float64 and F._real_of(np.bool_) == np.float64
    assert F._cplx_of(np.float32) == np.complex64 and F._cplx_of(np.int8) == np.complex128


def test_frequencies_and_fast_lengths():
    for n in (1, 2, 3, 8, 9, 1000, 1001):
        for d in (1.0, 0.1, 3):
            assert np.array_equal(F.fftfreq(n, d), np.fft.fftfreq(n, d))
            assert np.array_equal(F.rfftfreq(n, d), np.fft.rfftfreq(n, d))
    with pytest.raises(ValueError):
        F.rfftfreq(0)
    for t in list(range(0, 300)) + [1000, 1021, 15015, 65537, 1000003]:
        for real in (False, True):
            assert F.next_fast_len(t, real) == scipy.fft.next_fast_len(t, real), (t, real)
    with pytest.raises(ValueError):
        F.next_fast_len(-1)


def test_prev_fast_len_matches_scipy():
    """prev_fast_len is host-only integer code: bit-exact against scipy.fft.prev_fast_len (SciPy >= 1.14)."""
    import random

    import scipy.fft

    from rocket_fft_b200 import fft as F

    if not hasattr(scipy.fft, "prev_fast_len"):
        pytest.skip("scipy.fft.prev_fast_len not available")
    random.seed(3)
    targets = list(range(0, 4100)) + [random.randrange(1, 10**8) for _ in range(2000)] + [2**31 - 1, 2**31 + 1, 2**40 + 1]
    for t in targets:
        for real in (False, True):
            assert F.prev_fast_len(t, real) == scipy.fft.prev_fast_len(t, real), (t, real)
    with pytest.raises(ValueError):
        F.prev_fast_len(-1)
    # never above the target, never below the previous power of two, consistent with next_fast_len
    for t in (13, 1000003, 15015):
        p = F.prev_fast_len(t)
        assert p <= t and F.next_fast_len(p) == p


class _ReferenceBackend:
    """The low-level calls of fft.py served by the compiled, unmodified reference (oracle/_ref) instead of the CUDA
    library: lets the whole numpy/scipy-style host layer (padding, cropping, norm factors, axis handling, dtype
    promotion, type inversion) be checked against SciPy on a machine without a GPU."""

    def __init__(self, ref):
        self.ref = ref

    def c2c(self, a, o, axes, fwd, fct):
        return self.ref.c2c(a, o, list(axes), fwd, fct, 1)

    def c2c_sym(self, a, o, axes, fwd, fct):
        return self.ref.c2c_sym(a, o, list(axes), fwd, fct, 1)

    def r2c(self, a, o, axes, fwd, fct):
        return self.ref.r2c(a, o, list(axes), fwd, fct, 1)

    def c2r(self, a, o, axes, fwd, fct):
        return self.ref.c2r(a, o, list(axes), fwd, fct, 1)

    def dct(self, a, o, axes, type, fct, ortho):
        return self.ref.dct(a, o, list(axes), type, fct, ortho, 1)

    def dst(self, a, o, axes, type, fct, ortho, dst_ortho=None):
        # (the reference has one DST-II/III ortho scaling -- its own; the SciPy one exists only in the CUDA library)
        return self.ref.dst(a, o, list(axes), type, fct, ortho, 1)


def test_host_layer_on_top_of_the_reference_library_matches_scipy(monkeypatch):
    import parity

    ref = parity.reflib()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    monkeypatch.setattr(F, "_ll", _ReferenceBackend(ref))

    def close(a, b, tol=1e-11):
        assert a.shape == b.shape, (a.shape, b.shape)
        assert parity.l2err(a, b) < tol

    rng = np.random.default_rng(21)
    x = rng.standard_normal((5, 12, 9)) + 1j * rng.standard_normal((5, 12, 9))
    xr = rng.standard_normal((5, 12, 16))
    for norm in (None, "backward", "ortho", "forward"):
        for n in (None, 7, 12, 20):
            for axis in (-1, 0, 1):
                close(F.fft(x, n, axis, norm), scipy.fft.fft(x, n, axis, norm))
                close(F.ifft(x, n, axis, norm), scipy.fft.ifft(x, n, axis, norm))
                close(F.fft(xr, n, axis, norm), scipy.fft.fft(xr, n, axis, norm))
                close(F.rfft(xr, n, axis, norm), scipy.fft.rfft(xr, n, axis, norm))
                close(F.irfft(x, n, axis, norm), scipy.fft.irfft(x, n, axis, norm))
                close(F.hfft(x, n, axis, norm), scipy.fft.hfft(x, n, axis, norm))
                close(F.ihfft(xr, n, axis, norm), scipy.fft.ihfft(xr, n, axis, norm))
        for s, axes in ((None, (-2, -1)), ((10, 14), (-2, -1)), ((7, 9), (0, 2)), ((16, 6), (1, 0))):
            close(F.fft2(x, s, axes, norm), scipy.fft.fft2(x, s, axes, norm=norm))
            close(F.ifft2(x, s, axes, norm), scipy.fft.ifft2(x, s, axes, norm=norm))
            close(F.rfft2(xr, s, axes, norm), scipy.fft.rfft2(xr, s, axes, norm=norm))
            close(F.irfft2(x, s, axes, norm), scipy.fft.irfft2(x, s, axes, norm=norm))
            close(F.hfft2(x, s, axes, norm), scipy.fft.hfft2(x, s, axes, norm=norm))
            close(F.ihfft2(xr, s, axes, norm), scipy.fft.ihfft2(xr, s, axes, norm=norm))
        for s, axes in ((None, None), ((4, 12, 14), None), ((6, 11), (0, 1)), (None, (2, 0, 1))):
            close(F.fftn(x, s, axes, norm), scipy.fft.fftn(x, s, axes, norm=norm))
            close(F.ifftn(x, s, axes, norm), scipy.fft.ifftn(x, s, axes, norm=norm))
            close(F.rfftn(xr, s, axes, norm), scipy.fft.rfftn(xr, s, axes, norm=norm))
            close(F.irfftn(x, s, axes, norm), scipy.fft.irfftn(x, s, axes, norm=norm))
            close(F.hfftn(x, s, axes, norm), scipy.fft.hfftn(x, s, axes, norm=norm))
            close(F.ihfftn(xr, s, axes, norm), scipy.fft.ihfftn(xr, s, axes, norm=norm))
        for type in (1, 2, 3, 4):
            for n in (None, 9, 20):
                close(F.dct(xr, type, n, 1, norm), scipy.fft.dct(xr, type, n, 1, norm))
                close(F.idct(xr, type, n, 1, norm), scipy.fft.idct(xr, type, n, 1, norm))
            close(F.dctn(xr, type, None, (0, 2), norm), scipy.fft.dctn(xr, type, None, (0, 2), norm))
            close(F.idctn(xr, type, (6, 10), (0, 1), norm), scipy.fft.idctn(xr, type, (6, 10), (0, 1), norm))
        # DST: the default norm only, as in the reference's own comparison grid (tests/test_scipy_compare.py:365-367)
    for type in (1, 2, 3, 4):
        close(F.dst(xr, type, None, 1), scipy.fft.dst(xr, type, None, 1))
        close(F.idst(xr, type, 9, 0), scipy.fft.idst(xr, type, 9, 0))
        close(F.dstn(xr, type, None, (0, 2)), scipy.fft.dstn(xr, type, None, (0, 2)))
        close(F.idstn(xr, type, (6, 10), (0, 1)), scipy.fft.idstn(xr, type, (6, 10), (0, 1)))
    assert F.hfftn(x.astype(np.complex64)).dtype == np.float32


def test_numpy_like_and_scipy_like_duplicate_axes_policy(monkeypatch):
    """rocket_fft.numpy_like() / scipy_like() (O:325-339, 380-405): SciPy rejects duplicate axes, NumPy transforms the
    axis twice.  Checked on CPU with the compiled reference as the backend."""
    import parity
    import rocket_fft_b200 as R

    x = np.arange(24.0).reshape(4, 6) + 1j
    F.scipy_like()
    try:
        with pytest.raises(ValueError, match="unique"):
            F._shape_axes(x, None, (0, 0), True)
        with pytest.raises(ValueError, match="unique"):
            F._shape_axes(x, (4, 4), (1, -1), True)
        with pytest.raises(ValueError, match="more axes"):
            F._shape_axes(x, (2, 3, 4), None, True)
        assert F._shape_axes(x, None, (1, 0), True) == ([6, 4], [1, 0])
        R.numpy_like()
        assert F._shape_axes(x, None, (0, 0), True) == ([4, 4], [0, 0])
        ref = parity.reflib()
        if ref is not None:
            monkeypatch.setattr(F, "_ll", _ReferenceBackend(ref))
            got = F.fft2(x, axes=(0, 0))
            assert parity.l2err(got, np.fft.fft2(x, axes=(0, 0))) < 1e-12
    finally:
        R.scipy_like()
    assert R.get_workers() == 1
    R.set_workers(1)
    with pytest.raises(ValueError):
        R.set_workers(0)
    with pytest.raises(ValueError):
        R.set_workers(10**6)


def test_scipy_keyword_arguments_are_accepted_and_validated(monkeypatch):
    """overwrite_x / workers / plan of the scipy.fft signatures (validated like O:555-572, otherwise without effect)."""
    import parity

    ref = parity.reflib()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    monkeypatch.setattr(F, "_ll", _ReferenceBackend(ref))
    x = np.arange(12.0).reshape(3, 4)
    keep = x.copy()
    a = F.fft(x, None, -1, None, True, 2)            # positional like scipy.fft.fft(x, n, axis, norm, overwrite_x, workers)
    assert parity.l2err(a, scipy.fft.fft(x)) < 1e-12 and np.array_equal(x, keep)
    assert parity.l2err(F.rfft2(x, workers=-1, overwrite_x=True), scipy.fft.rfft2(x)) < 1e-12
    assert parity.l2err(F.dctn(x, 2, workers=1, overwrite_x=False), scipy.fft.dctn(x, 2)) < 1e-12
    with pytest.raises(ValueError, match="zero"):
        F.ifft(x, workers=0)
    with pytest.raises(ValueError, match="range"):
        F.fftn(x, workers=-(10**6))
    with pytest.raises(NotImplementedError):
        F.fft(x, plan=object())
