"""GPU tests of the numpy.fft / scipy.fft style layer (rocket_fft_b200.fft) against NumPy and
SciPy -- the grid of the reference's tests/test_numpy_compare.py and tests/test_scipy_compare.py
(norm x n/s x axes; all DCT/DST types; DST compared for the default norm only, as there)."""
import numpy as np
import pytest
import scipy.fft

import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    from rocket_fft_b200 import fft

    return fft


def close(a, b, tol=1e-11):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert parity.l2err(a, b) < tol


NORMS = (None, "backward", "ortho", "forward")


def test_complex_family(F):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((6, 42, 10)) + 1j * rng.standard_normal((6, 42, 10))
    xr = rng.standard_normal((6, 42, 10))
    for norm in NORMS:
        for n in (None, 8, 42, 55):
            for axis in (-1, 0, 1):
                close(F.fft(x, n, axis, norm), scipy.fft.fft(x, n, axis, norm))
                close(F.ifft(x, n, axis, norm), scipy.fft.ifft(x, n, axis, norm))
                close(F.fft(xr, n, axis, norm), scipy.fft.fft(xr, n, axis, norm))
                close(F.rfft(xr, n, axis, norm), scipy.fft.rfft(xr, n, axis, norm))
                close(F.irfft(x, n, axis, norm), scipy.fft.irfft(x, n, axis, norm))
                close(F.hfft(x, n, axis, norm), scipy.fft.hfft(x, n, axis, norm))
                close(F.ihfft(xr, n, axis, norm), scipy.fft.ihfft(xr, n, axis, norm))
        for s, axes in ((None, None), ((8, 12), (0, 1)), ((7, 50, 3), None), ((5,), (1,)), (None, (2, 0))):
            close(F.fftn(x, s, axes, norm), scipy.fft.fftn(x, s, axes, norm))
            close(F.ifftn(x, s, axes, norm), scipy.fft.ifftn(x, s, axes, norm))
            close(F.fftn(xr, s, axes, norm), scipy.fft.fftn(xr, s, axes, norm))
            close(F.rfftn(xr, s, axes, norm), scipy.fft.rfftn(xr, s, axes, norm))
            close(F.irfftn(x, s, axes, norm), scipy.fft.irfftn(x, s, axes, norm))
        close(F.fft2(x, norm=norm), np.fft.fft2(x, norm=norm))
        close(F.ifft2(x, norm=norm), np.fft.ifft2(x, norm=norm))
        close(F.rfft2(xr, norm=norm), np.fft.rfft2(xr, norm=norm))
        close(F.irfft2(x, norm=norm), np.fft.irfft2(x, norm=norm))
    x32 = x.astype(np.complex64)
    assert F.fft(x32).dtype == np.complex64 and F.rfft(xr.astype(np.float32)).dtype == np.complex64
    assert F.fft(np.arange(8)).dtype == np.complex128
    close(F.fft(np.arange(8)), np.fft.fft(np.arange(8)))


def test_dct_dst_family(F):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((9, 16, 5))
    z = x + 1j * rng.standard_normal(x.shape)
    for t in (1, 2, 3, 4):
        for norm in NORMS:
            for orth in (None, False, True):
                for n, axis in ((None, -1), (12, 1), (20, 0)):
                    close(F.dct(x, t, n, axis, norm, orth), scipy.fft.dct(x, t, n, axis, norm, orthogonalize=orth))
                    close(F.idct(x, t, n, axis, norm, orth), scipy.fft.idct(x, t, n, axis, norm, orthogonalize=orth))
            close(F.dctn(x, t, None, (0, 1), norm), scipy.fft.dctn(x, t, None, (0, 1), norm))
            close(F.idctn(x, t, (8, 8), (1, 2), norm), scipy.fft.idctn(x, t, (8, 8), (1, 2), norm))
        close(F.dct(z, t), scipy.fft.dct(z, t))
        # DST: default norm / orthogonalize only (tests/test_scipy_compare.py:365-367)
        for n, axis in ((None, -1), (12, 1), (20, 0)):
            close(F.dst(x, t, n, axis), scipy.fft.dst(x, t, n, axis))
            close(F.idst(x, t, n, axis), scipy.fft.idst(x, t, n, axis))
        close(F.dstn(x, t, None, (0, 2)), scipy.fft.dstn(x, t, None, (0, 2)))
        close(F.idstn(x, t), scipy.fft.idstn(x, t))
        for norm in ("ortho", "forward"):
            close(F.dst(x, t, norm=norm, orthogonalize=False), scipy.fft.dst(x, t, norm=norm, orthogonalize=False))


def test_next_fast_len_and_errors(F):
    for n in (1, 8, 16, 31, 1000003):
        for real in (False, True):
            assert F.next_fast_len(n, real) == scipy.fft.next_fast_len(n, real)
    with pytest.raises(ValueError):
        F.fft(np.zeros(4), norm="bogus")
    with pytest.raises(ValueError):
        F.fftn(np.zeros((4, 4)), s=(2,), axes=(0, 1))
    with pytest.raises(ValueError):
        F.dct(np.zeros(4), type=5)
    with pytest.raises(TypeError):
        F.rfft(np.zeros(4, dtype=np.complex128))


def test_torch_device_tensors(F):
    import torch

    x = torch.randn(16, 1000, dtype=torch.float32, device="cuda")
    X = F.rfft2(x)
    assert X.is_cuda and X.dtype == torch.complex64
    ref = torch.fft.rfft2(x)
    assert float(torch.linalg.vector_norm(torch.view_as_real(X - ref)) / torch.linalg.vector_norm(torch.view_as_real(ref))) < 2e-4
    y = F.irfft2(X, s=(16, 1000))
    assert float(torch.linalg.vector_norm(y - x) / torch.linalg.vector_norm(x)) < 2e-4
    d = F.dctn(x.double(), axes=(1,))
    want = scipy.fft.dctn(x.double().cpu().numpy(), axes=(1,))
    assert parity.l2err(d.cpu().numpy(), want) < 1e-11
