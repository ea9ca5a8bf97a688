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
    ref = scipy.fft.rfft2(x.cpu().numpy())
    assert parity.l2err(X.cpu().numpy(), ref) < 2e-4
    y = F.irfft2(X, s=(16, 1000))
    assert float(torch.linalg.vector_norm(y - x) / torch.linalg.vector_norm(x)) < 2e-4
    d = F.dctn(x.double(), axes=(1,))
    want = scipy.fft.dctn(x.double().cpu().numpy(), axes=(1,))
    assert parity.l2err(d.cpu().numpy(), want) < 1e-11


def test_fused_padding_and_cropping_on_device(F):
    """`n` / `s` on device tensors: zero-padding happens in the first load of every line (rfb200_*_pad),
    cropping is a strided view -- same results as SciPy's pad-then-transform (reference: O:575-609)."""
    import torch

    import rocket_fft_b200 as R

    rng = np.random.default_rng(5)
    for dt, tolv in ((np.complex128, 1e-11), (np.complex64, 2e-5)):
        rdt = np.float64 if dt == np.complex128 else np.float32
        x = (rng.standard_normal((6, 42, 10)) + 1j * rng.standard_normal((6, 42, 10))).astype(dt)
        xr = rng.standard_normal((6, 42, 10)).astype(rdt)
        d, dr = torch.from_numpy(x).cuda(), torch.from_numpy(xr).cuda()
        for norm in (None, "ortho"):
            for n in (8, 42, 55, 64, 100, 128, 4096):
                for axis in (-1, 0, 1):
                    close(F.fft(d, n, axis, norm).cpu().numpy(), scipy.fft.fft(x, n, axis, norm), tolv)
                    close(F.ifft(d, n, axis, norm).cpu().numpy(), scipy.fft.ifft(x, n, axis, norm), tolv)
                    close(F.rfft(dr, n, axis, norm).cpu().numpy(), scipy.fft.rfft(xr, n, axis, norm), tolv)
                    close(F.irfft(d, n, axis, norm).cpu().numpy(), scipy.fft.irfft(x, n, axis, norm), tolv)
                    close(F.hfft(d, n, axis, norm).cpu().numpy(), scipy.fft.hfft(x, n, axis, norm), tolv)
                    close(F.ihfft(dr, n, axis, norm).cpu().numpy(), scipy.fft.ihfft(xr, n, axis, norm), tolv)
            for s, axes in (((8, 12), (0, 1)), ((7, 50, 3), None), ((9, 64, 16), None), ((64,), (1,)), ((12, 3), (2, 0)),
                            ((16, 60), (0, 1)), ((3, 128), (1, 2))):
                close(F.fftn(d, s, axes, norm).cpu().numpy(), scipy.fft.fftn(x, s, axes, norm), tolv)
                close(F.ifftn(d, s, axes, norm).cpu().numpy(), scipy.fft.ifftn(x, s, axes, norm), tolv)
                close(F.rfftn(dr, s, axes, norm).cpu().numpy(), scipy.fft.rfftn(xr, s, axes, norm), tolv)
                close(F.irfftn(d, s, axes, norm).cpu().numpy(), scipy.fft.irfftn(x, s, axes, norm), tolv)
    # a padded transform launches no copy / fill kernels: one line kernel per axis
    big = torch.randn(64, 1000, dtype=torch.complex64, device="cuda")
    R.launch_count_reset()
    F.fft(big, 2048, -1)
    assert R.launch_count() == 1
    R.launch_count_reset()
    F.rfft(big.real.contiguous(), 2048, -1)   # packed half-length transform with the padding in its load
    assert R.launch_count() == 1
    # long padded lines (four-step / gather path) and a non-contiguous input
    x = (rng.standard_normal((3, 70000)) + 1j * rng.standard_normal((3, 70000))).astype(np.complex128)
    d = torch.from_numpy(x).cuda()
    close(F.fft(d, 131072, -1).cpu().numpy(), scipy.fft.fft(x, 131072, -1), 1e-11)
    close(F.fft(d.t(), 100, 1).cpu().numpy(), scipy.fft.fft(x.T, 100, 1), 1e-11)
    # shapes that differ along an untransformed axis are rejected
    with pytest.raises(R.TransformError):
        R.c2c_pad(d, torch.empty(4, 70000, dtype=torch.complex128, device="cuda"), [1], True, 1.0)


def test_shift_roll_and_frequencies(F):
    """fftshift / ifftshift / roll are one pass of the rotation kernel: bit-exact against NumPy
    (reference: rocket_fft/overloads.py:752-855, 1221-1310)."""
    import torch

    import rocket_fft_b200 as R

    rng = np.random.default_rng(6)
    for shape in ((1,), (7,), (8,), (5, 6), (4, 1, 9), (3, 4, 5, 6), (2, 3, 2, 3, 2, 3, 2, 3, 2)):
        for dt in (np.float32, np.float64, np.complex64, np.complex128, np.int32, np.int64, np.float16, np.int8):
            x = (rng.standard_normal(shape) * 100).astype(dt)
            if np.dtype(dt).kind == "c":
                x = x + 1j * (rng.standard_normal(shape) * 100).astype(dt)
            d = torch.from_numpy(x).cuda()
            nd = len(shape)
            for axes in (None, 0, -1, tuple(range(nd)), (0, 0), (nd - 1, 0)):
                assert np.array_equal(F.fftshift(d, axes).cpu().numpy(), np.fft.fftshift(x, axes))
                assert np.array_equal(F.ifftshift(d, axes).cpu().numpy(), np.fft.ifftshift(x, axes))
            assert np.array_equal(F.fftshift(x), np.fft.fftshift(x))      # host arrays are staged
            for shift, axis in ((3, None), (-2, 0), ((1, -5), (0, nd - 1)), (10 ** 6 + 1, -1)):
                assert np.array_equal(F.roll(d, shift, axis).cpu().numpy(), np.roll(x, shift, axis))
    # strided input / output through the low-level call, large array (> 2^32 bytes is peeled on the host)
    x = torch.randn(512, 1000, device="cuda")
    out = torch.empty(1000, 512, device="cuda").t()
    R.roll(x.t().contiguous().t(), out, [256, 500])
    assert torch.equal(out, torch.roll(x, (256, 500), (0, 1)))
    big = torch.arange(1 << 26, dtype=torch.float32, device="cuda").reshape(8192, 8192)
    assert torch.equal(F.fftshift(big), torch.fft.fftshift(big))
    with pytest.raises(ValueError):
        F.fftshift(big, axes=2)
    for n in (1, 2, 7, 8, 1001):
        for dd in (1.0, 0.25):
            assert np.array_equal(F.fftfreq(n, dd), np.fft.fftfreq(n, dd))
            assert np.array_equal(F.rfftfreq(n, dd), np.fft.rfftfreq(n, dd))
    assert F.fftfreq(8, device="cuda").is_cuda
    with pytest.raises(ValueError):
        F.fftfreq(0)


def test_multi_axis_hermitian_family(F):
    """scipy.fft.hfft2 / ihfft2 / hfftn / ihfftn (reference: rocket_fft/overloads.py:1516-1570): norm x s x axes."""
    rng = np.random.default_rng(7)
    x = rng.standard_normal((5, 12, 9)) + 1j * rng.standard_normal((5, 12, 9))
    xr = rng.standard_normal((5, 12, 16))
    for norm in NORMS:
        for s, axes in ((None, (-2, -1)), ((10, 14), (-2, -1)), ((7, 9), (0, 2)), ((16, 6), (1, 0))):
            close(F.hfft2(x, s, axes, norm), scipy.fft.hfft2(x, s, axes, norm=norm))
            close(F.ihfft2(xr, s, axes, norm), scipy.fft.ihfft2(xr, s, axes, norm=norm))
        for s, axes in ((None, None), ((4, 12, 14), None), ((6, 11), (0, 1)), (None, (2, 0, 1))):
            close(F.hfftn(x, s, axes, norm), scipy.fft.hfftn(x, s, axes, norm=norm))
            close(F.ihfftn(xr, s, axes, norm), scipy.fft.ihfftn(xr, s, axes, norm=norm))
    # single precision stays single
    x32 = x.astype(np.complex64)
    out = F.hfftn(x32)
    assert out.dtype == np.float32
    close(out, scipy.fft.hfftn(x32), 1e-5)
