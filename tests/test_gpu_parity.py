"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the C ABI
of librocketfft_b200.so -- the numba_* symbols with host arrays (H2D/D2H staged inside)
or the rfb200_* device entry points with torch CUDA tensors -- and is compared with the
compiled reference (oracle/_ref, when it travelled with the snapshot), the NumPy oracle
and the committed golden vectors.  Tolerance: rel-L2 <= 1e-5*log2(n) (fp32),
1e-13*log2(n) (fp64), n = product of the transformed lengths (north-star bound)."""
import itertools
import os

import numpy as np
import pytest

import parity
from oracle import pocketfft_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import rocket_fft_b200 as r

    return r


def trusted():
    """The checker: the compiled reference if present, else the NumPy oracle."""
    ref = parity.reflib()
    return ref if ref is not None else O


def cplx(rng, shape, dt):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)


CD = {np.float32: np.complex64, np.float64: np.complex128}


def check(got, want, dt, n, what=""):
    e = parity.l2err(got, want)
    assert e <= parity.tol(dt, n), (what, e, parity.tol(dt, n))


# ------------------------------------------------------------------------------------
def test_golden_vectors(R):
    d, cases = parity.golden()
    for i, c in enumerate(cases):
        ain = d[f"c{i}_in"]
        want = d[f"c{i}_out"]
        got = np.zeros_like(want)
        parity.call(R, c, ain.copy(), got)
        shape = want.shape if c["op"] == "c2r" else ain.shape
        n = parity.tlen(c, shape)
        if c["op"] in ("dct", "dst") and c["type"] == 1:
            n = 2 * n
        check(got, want, want.dtype, n, (i, c))


def test_readme_example(R):
    d, _ = parity.golden()
    out = np.empty(8, dtype=np.complex128)
    R.c2c(d["readme_in"], out, [0], True, 1.0)
    check(out, d["readme_out"], np.complex128, 8)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_c2c_every_length_to_600(R, dt):
    T = trusted()
    rng = np.random.default_rng(1)
    cdt = CD[dt]
    for n in range(1, 601):
        x = cplx(rng, (3, n), cdt)
        for fwd in (True, False):
            a, b = np.empty_like(x), np.empty_like(x)
            R.c2c(x, a, [1], fwd, 1.0)
            T.c2c(x, b, [1], fwd, 1.0)
            check(a, b, dt, n, (n, fwd))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_c2c_selected_lengths(R, dt):
    T = trusted()
    rng = np.random.default_rng(2)
    cdt = CD[dt]
    lens = [601, 625, 729, 1000, 1021, 1024, 1331, 2011, 2047, 2048, 2187, 4096, 5400, 7776, 8192, 15015, 16384,
            32768, 65536, 65537, 100003, 3 * 5 * 7 * 11 * 13 * 4, 2**18, 2**20, 1000003, 2000376]
    for n in lens:
        rows = 2 if n > 100000 else 5
        x = cplx(rng, (rows, n), cdt)
        a, b = np.empty_like(x), np.empty_like(x)
        R.c2c(x, a, [1], True, 1.0)
        T.c2c(x, b, [1], True, 1.0)
        check(a, b, dt, n, n)
        R.c2c(a, a, [1], False, 1.0 / n)  # in place, backward: round trip
        check(a, x, dt, n, ("roundtrip", n))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_c2c_shapes_strides_axes(R, dt):
    T = trusted()
    rng = np.random.default_rng(3)
    cdt = CD[dt]
    for shp in ((10,), (127,), (128, 128), (128, 129), (1, 129), (129, 1), (32, 17, 39), (4, 3, 5, 6)):
        nd = len(shp)
        x = cplx(rng, shp, cdt)
        axes_list = [list(range(nd)), [nd - 1], [0]]
        if nd >= 2:
            axes_list += [list(p) for p in itertools.permutations(range(nd))][:6] + [[0, 0], [nd - 1, nd - 1, 0]]
        for axes in axes_list:
            n = int(np.prod([shp[a] for a in set(axes)]))
            a, b = np.empty_like(x), np.empty_like(x)
            R.c2c(x, a, axes, True, 0.5)
            T.c2c(x, b, axes, True, 0.5)
            check(a, b, dt, n, (shp, axes))
    # views: F-order, reversed, sliced, transposed output
    base = cplx(rng, (24, 20, 18), cdt)
    views = [np.asfortranarray(base), base[::-1, :, ::-1], base[::2, ::3, :], base.transpose(2, 0, 1), base[3:, 1:, 2:]]
    for v in views:
        for axes in ([0], [1], [2], [0, 1, 2], [2, 0]):
            n = int(np.prod([v.shape[a] for a in set(axes)]))
            a = np.empty(v.shape, dtype=cdt)
            b = np.empty(v.shape, dtype=cdt)
            R.c2c(v, a, axes, False, 1.0)
            T.c2c(v, b, axes, False, 1.0)
            check(a, b, dt, n, (v.strides, axes))
            af = np.asfortranarray(np.zeros(v.shape, dtype=cdt))
            R.c2c(v, af, axes, False, 1.0)
            check(af, b, dt, n, ("F-out", v.strides, axes))
    # in place
    x = cplx(rng, (64, 48), cdt)
    want = np.empty_like(x)
    T.c2c(x, want, [0, 1], True, 1.0)
    assert R.c2c(x, x, [0, 1], True, 1.0) is x
    check(x, want, dt, 64 * 48)
    # zero-sized dims are a silent no-op
    z = np.zeros((0, 5), dtype=cdt)
    R.c2c(z, z, [1], True, 1.0)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_real_transforms_lengths(R, dt):
    T = trusted()
    rng = np.random.default_rng(4)
    cdt = CD[dt]
    for n in list(range(1, 131)) + [255, 256, 257, 1000, 1021, 2048, 4096, 16384, 30000, 65536, 100003]:
        x = rng.standard_normal((2, n)).astype(dt)
        z = cplx(rng, (2, n // 2 + 1), cdt)
        for fwd in (True, False):
            a = np.zeros((2, n // 2 + 1), dtype=cdt)
            b = np.zeros_like(a)
            R.r2c(x, a, [1], fwd, 1.0)
            T.r2c(x, b, [1], fwd, 1.0)
            check(a, b, dt, n, ("r2c", n, fwd))
            a, b = np.empty_like(x), np.empty_like(x)
            R.c2r(z, a, [1], fwd, 1.0)
            T.c2r(z, b, [1], fwd, 1.0)
            check(a, b, dt, n, ("c2r", n, fwd))
            a = np.empty((2, n), dtype=cdt)
            b = np.empty_like(a)
            R.c2c_sym(x, a, [1], fwd, 1.0)
            T.c2c_sym(x, b, [1], fwd, 1.0)
            check(a, b, dt, n, ("c2c_sym", n, fwd))
            if n > 5000:
                continue
            for r2h in (True, False):
                a, b = np.empty_like(x), np.empty_like(x)
                R.r2r_fftpack(x, a, [1], r2h, fwd, 1.0)
                T.r2r_fftpack(x, b, [1], r2h, fwd, 1.0)
                check(a, b, dt, n, ("fftpack", n, r2h, fwd))
        a, b = np.empty_like(x), np.empty_like(x)
        R.r2r_separable_hartley(x, a, [1], 1.0)
        T.r2r_separable_hartley(x, b, [1], 1.0)
        check(a, b, dt, n, ("hartley", n))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_real_transforms_nd(R, dt):
    T = trusted()
    rng = np.random.default_rng(5)
    cdt = CD[dt]
    for shp in ((128, 128), (128, 129), (1, 129), (129, 1), (32, 17, 39), (12, 10, 9)):
        nd = len(shp)
        x = rng.standard_normal(shp).astype(dt)
        for axes in ([list(range(nd)), [nd - 1], [0], [nd - 1, 0], [1, 0]]):
            n = int(np.prod([shp[a] for a in set(axes)]))
            oshp = list(shp)
            oshp[axes[-1]] = shp[axes[-1]] // 2 + 1
            a = np.zeros(oshp, dtype=cdt)
            b = np.zeros(oshp, dtype=cdt)
            R.r2c(x, a, axes, True, 1.0)
            T.r2c(x, b, axes, True, 1.0)
            check(a, b, dt, n, ("r2c", shp, axes))
            zin = cplx(rng, oshp, cdt)
            a2, b2 = np.empty_like(x), np.empty_like(x)
            R.c2r(zin, a2, axes, False, 1.0 / n)
            T.c2r(zin, b2, axes, False, 1.0 / n)
            check(a2, b2, dt, n, ("c2r", shp, axes))
            a3 = np.empty(shp, dtype=cdt)
            b3 = np.empty(shp, dtype=cdt)
            R.c2c_sym(x, a3, axes, True, 1.0)
            T.c2c_sym(x, b3, axes, True, 1.0)
            check(a3, b3, dt, n, ("c2c_sym", shp, axes))
            a4, b4 = np.empty_like(x), np.empty_like(x)
            R.r2r_genuine_hartley(x, a4, axes, 1.0)
            T.r2r_genuine_hartley(x, b4, axes, 1.0)
            check(a4, b4, dt, n, ("genuine", shp, axes))
            R.r2r_separable_hartley(x, a4, axes, 1.0)
            T.r2r_separable_hartley(x, b4, axes, 1.0)
            check(a4, b4, dt, n, ("separable", shp, axes))
            R.r2r_fftpack(x, a4, axes, True, True, 1.0)
            T.r2r_fftpack(x, b4, axes, True, True, 1.0)
            check(a4, b4, dt, n, ("fftpack", shp, axes))
    # Hartley identities of the reference's own tests (tests/test_low_level_interface.py:90-167)
    x = (rng.random((32, 17, 39)) - 0.5).astype(dt)
    axes = [0, 1, 2]
    y = np.empty_like(x)
    R.r2r_genuine_hartley(x, y, axes, 1.0)
    v = np.fft.fftn(x.astype(np.complex128))
    check(y, v.real + v.imag, dt, x.size, "genuine == Re+Im fftn")
    v1 = x.copy()
    f = 1 / np.sqrt(x.size)
    assert R.r2r_genuine_hartley(R.r2r_genuine_hartley(v1, v1, axes, f), v1, axes, f) is v1
    check(v1, x, dt, x.size, "genuine twice in place")
    y2 = np.empty_like(x)
    R.r2r_separable_hartley(R.r2r_separable_hartley(x, y, axes, 1.0), y2, axes, 1.0 / x.size)
    check(y2, x, dt, x.size, "separable twice")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_dct_dst(R, dt):
    T = trusted()
    rng = np.random.default_rng(6)
    for n in list(range(1, 40)) + [64, 100, 101, 128, 255, 256, 1000, 1024, 2048, 4097]:
        x = rng.standard_normal((3, n)).astype(dt)
        for kind in ("dct", "dst"):
            for t in (1, 2, 3, 4):
                if kind == "dct" and t == 1 and n < 2:
                    continue
                for ortho in (False, True):
                    a, b = np.empty_like(x), np.empty_like(x)
                    getattr(R, kind)(x, a, [1], t, 1.0, ortho)
                    getattr(T, kind)(x, b, [1], t, 1.0, ortho)
                    check(a, b, dt, 2 * n + 2, (kind, t, n, ortho))
    x = rng.standard_normal((40, 24, 6)).astype(dt)
    for kind in ("dct", "dst"):
        for t in (1, 2, 3, 4):
            for axes in ([0, 1], [2], [1, 0, 2], [0, 0]):
                a, b = np.empty_like(x), np.empty_like(x)
                getattr(R, kind)(x, a, axes, t, 0.25, False)
                getattr(T, kind)(x, b, axes, t, 0.25, False)
                check(a, b, dt, 4 * x.size, (kind, t, axes))
    # strided power-of-two lines (fused kernel, line-fast tiles), both precisions, in place
    x = rng.standard_normal((64, 128, 24)).astype(dt)
    for kind in ("dct", "dst"):
        for t in (2, 3):
            for ortho in (False, True):
                for axes in ([0], [1], [0, 1], [1, 0]):
                    b = np.empty_like(x)
                    getattr(T, kind)(x, b, axes, t, 0.5, ortho)
                    a = x.copy()
                    getattr(R, kind)(a, a, axes, t, 0.5, ortho)
                    check(a, b, dt, 4 * 64 * 128, (kind, t, axes, ortho, "inplace strided"))
    # FFTW known answers (tests/test_scipy_testsuite.py:1192-1293)
    d, _ = parity.golden()
    for k in [k for k in d.files if k.startswith("fftw_")]:
        _, kind, t, N = k.split("_")
        xx = np.linspace(0, int(N) - 1, int(N)).astype(dt)
        y = np.empty_like(xx)
        getattr(R, kind)(xx, y, [0], int(t), 1.0, False)
        check(y, d[k], dt, 4 * int(N), k)


def test_config1_c2c_complex128_4096x4096(R):
    """BASELINE config 1 at full size."""
    T = trusted()
    rng = np.random.default_rng(0)
    x = cplx(rng, (4096, 4096), np.complex128)
    a, b = np.empty_like(x), np.empty_like(x)
    R.c2c(x, a, [1], True, 1.0)
    T.c2c(x, b, [1], True, 1.0, 8) if T is not O else T.c2c(x, b, [1], True, 1.0)
    check(a, b, np.float64, 4096, "cfg1")


def test_config4_nonpow2_and_bluestein(R):
    T = trusted()
    rng = np.random.default_rng(3)
    x = cplx(rng, (256, 15015), np.complex64)
    a, b = np.empty_like(x), np.empty_like(x)
    R.c2c(x, a, [1], True, 1.0)
    T.c2c(x, b, [1], True, 1.0)
    check(a, b, np.float32, 15015, "cfg4a")
    x = cplx(rng, (8, 1000003), np.complex64)
    a, b = np.empty_like(x), np.empty_like(x)
    R.c2c(x, a, [1], True, 1.0)
    T.c2c(x, b, [1], True, 1.0)
    check(a, b, np.float32, 1000003, "cfg4b")


def test_config2_reduced_and_config5_reduced(R):
    T = trusted()
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2048, 2048)).astype(np.float32)
    a = np.zeros((2048, 1025), dtype=np.complex64)
    b = np.zeros_like(a)
    R.r2c(x, a, [0, 1], True, 1.0)
    T.r2c(x, b, [0, 1], True, 1.0)
    check(a, b, np.float32, 2048 * 2048, "cfg2 r2c")
    y, y2 = np.empty_like(x), np.empty_like(x)
    R.c2r(a, y, [0, 1], False, 1.0 / x.size)
    check(y, x, np.float32, 2048 * 2048, "cfg2 roundtrip")
    x = rng.standard_normal((256, 256, 64))
    for kind in ("dct", "dst"):
        a, b = np.empty_like(x), np.empty_like(x)
        getattr(R, kind)(x, a, [0, 1], 2, 1.0, False)
        getattr(T, kind)(x, b, [0, 1], 2, 1.0, False)
        check(a, b, np.float64, 4 * 256 * 256, "cfg5 " + kind)


def test_device_arrays_and_full_size_properties(R):
    """rfb200_* entry points on torch CUDA tensors; full-size configs through
    size-independent properties (round trip, linearity, Parseval)."""
    import torch

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    # device path equals host path
    x = torch.randn(64, 1000, dtype=torch.complex128, device=dev, generator=g)
    y = torch.empty_like(x)
    R.c2c(x, y, [1], True, 1.0)
    torch.cuda.synchronize()
    xh = x.cpu().numpy()
    yh = np.empty_like(xh)
    R.c2c(xh, yh, [1], True, 1.0)
    assert parity.l2err(y.cpu().numpy(), yh) < 1e-15
    # config 2 full size: rfft2 -> irfft2 round trip, Parseval
    x = torch.randn(16384, 16384, dtype=torch.float32, device=dev, generator=g)
    X = torch.empty(16384, 8193, dtype=torch.complex64, device=dev)
    R.r2c(x, X, [0, 1], True, 1.0)
    e_time = float((x.double() ** 2).sum())
    w = torch.full((8193,), 2.0, dtype=torch.float64, device=dev)
    w[0] = 1.0
    w[-1] = 1.0
    e_freq = float(((X.real.double() ** 2 + X.imag.double() ** 2) * w).sum()) / x.numel()
    assert abs(e_freq - e_time) / e_time < 1e-5
    xr = torch.empty_like(x)
    R.c2r(X, xr, [0, 1], False, 1.0 / x.numel())
    err = float(torch.linalg.vector_norm((xr - x).double()) / torch.linalg.vector_norm(x.double()))
    assert err < parity.tol(np.float32, 2**28), err
    del x, X, xr
    # config 3 at 256^3 against the reference, then the full 1024^3 round trip
    T = trusted()
    vh = cplx(np.random.default_rng(2), (256, 256, 256), np.complex64)
    v = torch.from_numpy(vh).to(dev)
    V = torch.empty_like(v)
    R.c2c(v, V, [0, 1, 2], True, 1.0)
    want = np.empty_like(vh)
    T.c2c(vh, want, [0, 1, 2], True, 1.0)
    check(V.cpu().numpy(), want, np.float32, 2**24, "fftn 256^3")
    del v, V, want, vh
    v = torch.randn(1024, 1024, 1024, dtype=torch.complex64, device=dev, generator=g)
    V = torch.empty_like(v)
    R.c2c(v, V, [0, 1, 2], True, 1.0)
    R.c2c(V, V, [0, 1, 2], False, 1.0 / v.numel())
    err = float(torch.linalg.vector_norm((V - v)[:64].to(torch.complex128)) / torch.linalg.vector_norm(v[:64].to(torch.complex128)))
    assert err < parity.tol(np.float32, 2**30), err
    del v, V
    # config 5 full size: dct-II then dct-III inverts up to 1/(2N)^2
    x = torch.randn(2048, 2048, 64, dtype=torch.float64, device=dev, generator=g)
    y = torch.empty_like(x)
    R.dct(x, y, [0, 1], 2, 1.0, False)
    R.dct(y, y, [0, 1], 3, 1.0 / (4096.0 * 4096.0), False)
    err = float(torch.linalg.vector_norm(y - x) / torch.linalg.vector_norm(x))
    assert err < parity.tol(np.float64, 4096 * 4096), err


def test_baseline_configs_full_size_against_the_reference(R):
    """The BASELINE configs at FULL size, device result against the compiled reference (oracle/_ref, all host
    threads) on the same seeded input, under the north-star tolerance: cfg2 r2c AND c2r 16384^2, cfg5 dct AND dst
    (2048,2048,64), cfg4b Bluestein at 256 rows, and a 512^3 complex64 fftn (cfg3's code path; 1024^3 is 16 GiB of
    host memory per side and minutes of CPU time)."""
    import torch

    T = parity.reflib()
    if T is None:
        pytest.skip("oracle/_ref did not travel")
    nt = os.cpu_count() or 1
    dev = torch.device("cuda:0")

    def gpu(fn, ain, out_shape, out_dtype):
        d_in = torch.from_numpy(ain).to(dev)
        d_out = torch.empty(out_shape, dtype=out_dtype, device=dev)
        fn(d_in, d_out)
        torch.cuda.synchronize()
        res = d_out.cpu().numpy()
        del d_in, d_out
        torch.cuda.empty_cache()
        return res

    # ---- cfg2: rfft2 / irfft2 of a float32 16384 x 16384 image --------------------------------------------------------
    rng = np.random.default_rng(1)
    x = rng.standard_normal((16384, 16384), dtype=np.float32)
    want = np.empty((16384, 8193), dtype=np.complex64)
    T.r2c(x, want, [0, 1], True, 1.0, nt)
    got = gpu(lambda a, b: R.r2c(a, b, [0, 1], True, 1.0), x, (16384, 8193), torch.complex64)
    check(got, want, np.float32, 2**28, "cfg2 r2c full size")
    del got
    wantr = np.empty_like(x)
    T.c2r(want, wantr, [0, 1], False, 1.0 / 2**28, nt)
    gotr = gpu(lambda a, b: R.c2r(a, b, [0, 1], False, 1.0 / 2**28), want, (16384, 16384), torch.float32)
    check(gotr, wantr, np.float32, 2**28, "cfg2 c2r full size")
    check(gotr, x, np.float32, 2**28, "cfg2 round trip")
    del x, want, wantr, gotr
    # ---- cfg5: dctn / dstn type II, float64 (2048, 2048, 64), axes (0, 1) ------------------------------------------------
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2048, 2048, 64))
    for kind in ("dct", "dst"):
        want = np.empty_like(x)
        getattr(T, kind)(x, want, [0, 1], 2, 1.0, False, nt)
        got = gpu(lambda a, b: getattr(R, kind)(a, b, [0, 1], 2, 1.0, False), x, x.shape, torch.float64)
        check(got, want, np.float64, 4096 * 4096, "cfg5 full size " + kind)
        del want, got
    del x
    # ---- cfg4b: prime length 1000003 (Bluestein), 256 rows ---------------------------------------------------------------
    rng = np.random.default_rng(4)
    x = cplx(rng, (256, 1000003), np.complex64)
    want = np.empty_like(x)
    T.c2c(x, want, [1], True, 1.0, nt)
    got = gpu(lambda a, b: R.c2c(a, b, [1], True, 1.0), x, x.shape, torch.complex64)
    check(got, want, np.float32, 1000003, "cfg4b full size")
    del x, want, got
    # ---- cfg3's path: fftn of a complex64 512^3 volume -----------------------------------------------------------------
    rng = np.random.default_rng(2)
    x = cplx(rng, (512, 512, 512), np.complex64)
    want = np.empty_like(x)
    T.c2c(x, want, [0, 1, 2], True, 1.0, nt)
    got = gpu(lambda a, b: R.c2c(a, b, [0, 1, 2], True, 1.0), x, x.shape, torch.complex64)
    check(got, want, np.float32, 2**27, "fftn 512^3")
    back = gpu(lambda a, b: R.c2c(a, b, [2, 0, 1], False, 1.0 / 2**27), want, x.shape, torch.complex64)
    check(back, x, np.float32, 2**27, "ifftn 512^3 (axes permuted)")


def test_slab_pipeline_emulated_on_one_device(R):
    """The kernels of the slab-decomposed fftn (rocket_fft_b200.distributed.SlabFFTN, engine "fused") on ONE device,
    against the reference: the P ranks are played one after the other, each rank's axis-1 transform scatters its
    output blocks into the P receive buffers with rfb200_c2c_scatter -- the fused FFT + all-to-all push kernel; on a
    multi-GPU box those buffers are the peers' NVLink-mapped memory, here they are local -- and the axis-0 transforms
    follow.  (The real multi-rank run is tests/_slab_check.py under torchrun and bench.py --gpus N.)"""
    import torch

    T = trusted()
    dev = torch.device("cuda:0")
    for shape, ranks in (((128, 128, 128), (2, 8)), ((64, 1024, 96), (2, 4, 8)), ((16, 2048, 8), (4,))):
        n0, n1, n2 = shape
        fullh = cplx(np.random.default_rng(21), shape, np.complex64)
        want = np.empty_like(fullh)
        T.c2c(fullh, want, [0, 1, 2], True, 1.0)
        full = torch.from_numpy(fullh).to(dev)
        for P in ranks:
            recv = [torch.zeros(P, n0 // P, n1 // P, n2, dtype=torch.complex64, device=dev) for _ in range(P)]
            for g in range(P):
                x = full[g * (n0 // P):(g + 1) * (n0 // P)].clone()
                R.c2c(x, x, [2], True, 1.0)
                R.c2c_scatter(x, [recv[h][g] for h in range(P)], 1, True, 1.0)
            for h in range(P):
                y = recv[h].view(n0, n1 // P, n2)
                R.c2c(y, y, [0], True, 1.0)
                torch.cuda.synchronize()
                check(y.cpu().numpy(), want[:, h * (n1 // P):(h + 1) * (n1 // P)], np.float32, n0 * n1 * n2, ("slab emulation", shape, P, h))


def _dcst_fft_len(kind, typ, N):
    """Length of the complex transform a DCT / DST line of N points is folded into (ops.cu op_dcst)."""
    if typ == 1:
        return 2 * (N - 1) if kind == "dct" else 2 * (N + 1)
    if typ in (2, 3):
        return N
    return N // 2 if N % 2 == 0 else 2 * N


def _largest_prime(n):
    p, f = 1, 2
    while f * f <= n:
        while n % f == 0:
            p, n = f, n // f
        f += 1
    return max(p, n) if n > 1 else p


def test_dct_dst_one_launch_per_axis_and_no_work_area(R):
    """Every DCT / DST type, odd and even lengths, runs as ONE kernel launch per transformed axis (the type's reordering,
    extension and phase factors are the load / store stage of the line transform; reference: T_dct1 H:2918-2955, T_dst1
    H:2957-2985, T_dcst23 H:2987-3061, T_dcst4 H:3063-3163) -- against the reference, float64 and float32, strided axes,
    ortho on and off."""
    import torch

    T = trusted()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(31)
    for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        for shape, axes in (((6, 100), [1]), ((5, 101), [1]), ((1000, 12), [0]), ((33, 64, 10), [0, 1]), ((4, 2048), [1]),
                            ((3, 2049), [1]), ((2047, 5), [0]), ((7, 2), [1]), ((2, 3, 4), [0, 1, 2])):
            x = rng.standard_normal(shape).astype(dt)
            d_x = torch.from_numpy(x).to(dev)
            for kind in ("dct", "dst"):
                for typ in (1, 2, 3, 4):
                    for ortho in (False, True):
                        want = np.empty_like(x)
                        getattr(T, kind)(x, want, axes, typ, 0.5, ortho)
                        d_y = torch.empty_like(d_x)
                        R.launch_count_reset()
                        getattr(R, kind)(d_x, d_y, axes, typ, 0.5, ortho)
                        torch.cuda.synchronize()
                        if all(_largest_prime(_dcst_fft_len(kind, typ, shape[a])) <= 64 for a in axes):
                            # (a transform length with a prime factor above 64 -- e.g. DST-I of 100 points, 2 * 101 -- is a
                            # Bluestein line and keeps the work-area path)
                            assert R.launch_count() == len(axes), (kind, typ, shape, R.launch_count())
                        n = 1
                        for a in axes:
                            n *= 4 * shape[a]
                        check(d_y.cpu().numpy(), want, dt, n, (kind, typ, shape, ortho, dt))
                        # in place
                        d_z = d_x.clone()
                        getattr(R, kind)(d_z, d_z, axes, typ, 0.5, ortho)
                        torch.cuda.synchronize()
                        check(d_z.cpu().numpy(), want, dt, n, (kind, typ, shape, ortho, dt, "in place"))


def test_numba_njit_calls_run_on_the_gpu(R):
    """The reference's own usage pattern (tests/test_low_level_interface.py:30-51): low-level
    functions called from nogil @njit code on NumPy arrays, here bound to librocketfft_b200."""
    import numba as nb

    from rocket_fft_b200 import numba_api as napi

    @nb.njit(nogil=True)
    def jit_c2c(ain, aout, axes, forward, fct, nthreads):
        napi.c2c(ain, aout, axes, forward, fct, nthreads)
        return aout

    @nb.njit(nogil=True)
    def jit_hartley(ain, aout, axes, fct, nthreads):
        napi.r2r_genuine_hartley(ain, aout, axes, fct, nthreads)
        return aout

    rng = np.random.default_rng(8)
    x = cplx(rng, (32, 17, 39), np.complex128)
    want = np.fft.fftn(x)
    for axes_dtype in (np.int64, np.uint64, np.int32, np.uint8, np.float64):
        out = np.empty_like(x)
        axes = np.arange(3).astype(axes_dtype)
        R.launch_count_reset()
        jit_c2c(x, out, axes, True, 1.0, 1)
        assert R.launch_count() > 0
        check(out, want, np.float64, x.size, axes_dtype)
    v1 = x.real.copy()
    assert jit_c2c(x, x, np.arange(3), True, np.float32(1.0), np.int16(4)) is x
    check(x, want, np.float64, x.size, "in place")
    a = rng.random((128, 129)) - 0.5
    out = np.empty_like(a)
    jit_hartley(a, out, np.arange(2), 1.0, 1)
    w = np.fft.fftn(a)
    check(out, w.real + w.imag, np.float64, a.size, "hartley")


def test_reentrancy_many_python_threads(R):
    """The reference's tests/test_multithreading.py pattern: concurrent calls from 8 threads
    (here with the results actually checked)."""
    from concurrent.futures import ThreadPoolExecutor

    rng = np.random.default_rng(9)
    arrays = [cplx(rng, (4, 2**14), np.complex128) for _ in range(24)]
    reals = [rng.standard_normal((3, 1000)) for _ in range(24)]

    def work(i):
        x = arrays[i]
        out = np.empty_like(x)
        R.c2c(x, out, [1], True, 1.0)
        e1 = parity.l2err(out, np.fft.fft(x, axis=1))
        r = reals[i]
        d = np.empty_like(r)
        R.dct(r, d, [1], 2, 1.0, False)
        import scipy.fft

        e2 = parity.l2err(d, scipy.fft.dct(r, 2, axis=1))
        return max(e1, e2)

    with ThreadPoolExecutor(max_workers=8) as ex:
        errs = list(ex.map(work, range(24)))
    assert max(errs) < 1e-12, errs


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_bluestein_strided_and_batched(R, dt):
    """Prime lengths > 64 (chirp-z) along contiguous and strided axes, forward and backward,
    in place and with fct -- exercises the fused chirp/pad/multiply/truncate paths."""
    T = trusted()
    rng = np.random.default_rng(10)
    cdt = CD[dt]
    for shp, axes in (((67, 5, 3), [0]), ((131, 4), [0]), ((6, 4099), [1]), ((3, 2, 257), [2]), ((521, 33), [0, 1]),
                      ((2, 70001), [1]), ((70001, 2), [0])):
        x = cplx(rng, shp, cdt)
        n = int(np.prod([shp[a] for a in axes]))
        for fwd in (True, False):
            a, b = np.empty_like(x), np.empty_like(x)
            R.c2c(x, a, axes, fwd, 0.75)
            T.c2c(x, b, axes, fwd, 0.75)
            check(a, b, dt, n, (shp, axes, fwd))
        y = x.copy()
        R.c2c(y, y, axes, True, 1.0)
        T.c2c(x, b, axes, True, 1.0)
        check(y, b, dt, n, (shp, axes, "in place"))
    xr = rng.standard_normal((4, 4099)).astype(dt)
    a = np.zeros((4, 2050), dtype=cdt)
    b = np.zeros_like(a)
    R.r2c(xr, a, [1], True, 1.0)
    T.r2c(xr, b, [1], True, 1.0)
    check(a, b, dt, 4099, "r2c prime")


def test_runtime_specialised_kernels(tmp_path):
    """Smooth non-power-of-two lengths through the NVRTC-specialised kernel (csrc/jit.cu, RFB200_JIT=2 forces
    it for every eligible call), twice: the second process must take its cubins from the on-disk cache."""
    import subprocess
    import sys

    env = dict(os.environ, RFB200_JIT="2", RFB200_JIT_VERBOSE="1", RFB200_CACHE_DIR=str(tmp_path / "jitcache"))
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_jit_check.py")
    first = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=900)
    assert first.returncode == 0, first.stdout[-3000:] + first.stderr[-3000:]
    assert "rocketfft_b200: jit n=1000 float" in first.stderr, first.stderr[-2000:]
    assert "spill=" in first.stderr
    second = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=900)
    assert second.returncode == 0, second.stdout[-3000:] + second.stderr[-3000:]
    assert " from " in second.stderr and "spill=" not in second.stderr, second.stderr[-2000:]


def test_long_real_lines_two_transforms_per_thread(R):
    """16384-point float32 real lines: the kernel that runs two interleaved half-length transforms per thread
    (pow2_dual_kernel.cuh) on aligned contiguous lines, and the single-transform kernel on everything else --
    padded rows, shifted (4-byte aligned) views, strided outputs, 3-D batches, both directions, fct != 1,
    non-zero imaginary DC/Nyquist bins (ignored, H:3830) -- all against the reference."""
    import torch

    T = trusted()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    n = 16384
    nb = n // 2 + 1

    def run_r2c(xv, ov, fwd, fct):
        R.r2c(xv, ov, [xv.dim() - 1], fwd, fct)
        torch.cuda.synchronize()
        xh = np.ascontiguousarray(xv.cpu().numpy())
        want = np.zeros(xh.shape[:-1] + (nb,), dtype=np.complex64)
        T.r2c(xh, want, [xh.ndim - 1], fwd, fct)
        check(ov.cpu().numpy(), want, np.float32, n, ("r2c", tuple(xv.shape), tuple(xv.stride()), fwd))

    base = torch.from_numpy(rng.standard_normal((5, 3, n + 16)).astype(np.float32)).to(dev)
    for xv in (base[..., :n], base[..., 4 : n + 4], base[..., 1 : n + 1], base[:, 1, 8 : n + 8], base[2, 2, :n]):
        for fwd in (True, False):
            out = torch.zeros(tuple(xv.shape[:-1]) + (nb,), dtype=torch.complex64, device=dev)
            run_r2c(xv, out, fwd, 0.25)
    # strided output bins / padded output rows
    xv = base[0, :, :n]
    wide = torch.zeros(3, 2 * nb + 6, dtype=torch.complex64, device=dev)
    run_r2c(xv, wide[:, 0 : 2 * nb : 2], True, 1.0)
    run_r2c(xv, wide[:, 3 : 3 + nb], False, 2.0)
    # 2-D: rows by the two-transform kernel, columns by the four-step passes
    x2 = torch.from_numpy(rng.standard_normal((64, n)).astype(np.float32)).to(dev)
    o2 = torch.zeros(64, nb, dtype=torch.complex64, device=dev)
    R.r2c(x2, o2, [0, 1], True, 1.0)
    want = np.zeros((64, nb), dtype=np.complex64)
    T.r2c(x2.cpu().numpy(), want, [0, 1], True, 1.0)
    check(o2.cpu().numpy(), want, np.float32, 64 * n, "r2c 2-D")

    # ---- c2r ----
    zh = cplx(rng, (4, 3, nb + 5), np.complex64)
    z = torch.from_numpy(zh).to(dev)
    outb = torch.zeros(4, 3, n + 8, dtype=torch.float32, device=dev)
    for zv in (z[..., :nb], z[..., 2 : nb + 2], z[1, :, 1 : nb + 1]):
        for ov_full in (outb[..., :n], outb[..., 4 : n + 4], outb[..., 1 : n + 1]):
            ov = ov_full if zv.dim() == 3 else ov_full[1]
            for fwd in (True, False):
                R.c2r(zv, ov, [zv.dim() - 1], fwd, 0.5)
                torch.cuda.synchronize()
                zc = np.ascontiguousarray(zv.cpu().numpy())
                want = np.zeros(zc.shape[:-1] + (n,), dtype=np.float32)
                T.c2r(zc, want, [zc.ndim - 1], fwd, 0.5)
                check(ov.cpu().numpy(), want, np.float32, n, ("c2r", tuple(zv.stride()), tuple(ov.stride()), fwd))
    # round trip of a batch that fills the GPU several times over (every CTA slot, L2 prefetch of later tiles)
    xb = torch.from_numpy(rng.standard_normal((1200, n)).astype(np.float32)).to(dev)
    Xb = torch.empty(1200, nb, dtype=torch.complex64, device=dev)
    R.r2c(xb, Xb, [1], True, 1.0)
    want = np.zeros((1200, nb), dtype=np.complex64)
    T.r2c(xb.cpu().numpy(), want, [1], True, 1.0)
    check(Xb.cpu().numpy(), want, np.float32, n, "r2c batch of 1200 long lines")
    yb = torch.empty_like(xb)
    R.c2r(Xb, yb, [1], False, 1.0 / n)
    err = float(torch.linalg.vector_norm((yb - xb).double()) / torch.linalg.vector_norm(xb.double()))
    assert err < parity.tol(np.float32, n), err


def test_fused_fourstep_long_strided_lines(R):
    """16384-point complex64 lines along a strided axis of arrays larger than the L2 cache: both four-step passes run
    in one persistent kernel with the intermediate in an L2-resident scratch ring (fused4v2_kernel.cuh).  Widths that
    are not multiples of the strip / tile width, a batch of arrays, in place and out of place, both directions,
    fct != 1 -- against the reference on the host."""
    import torch

    T = trusted()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(12)
    n = 16384
    # RFB200_FUSE4 (read per call): unset / 1 = the fused kernel (default), 0 = two launches
    try:
        for mode in ("1", "0"):
            os.environ["RFB200_FUSE4"] = mode
            R.launch_trace(True)
            _fused_fourstep_cases(R, T, dev, rng, n)
            names = R.launch_trace_get()
            R.launch_trace(False)
            assert any("fused2" in k for k in names) == (mode == "1"), names[:6]
    finally:
        os.environ.pop("RFB200_FUSE4", None)
    _fused_fourstep_cases(R, T, dev, rng, n)


def _fused_fourstep_cases(R, T, dev, rng, n):
    import torch

    for shape, axis in (((n, 777), 0), ((2, n, 800), 1), ((n, 1025), 0), ((3, n, 545), 1)):
        xh = cplx(rng, shape, np.complex64)
        x = torch.from_numpy(xh).to(dev)
        for fwd, fct, inplace in ((True, 1.0, False), (False, 0.5, True)):
            want = np.empty_like(xh)
            T.c2c(xh, want, [axis], fwd, fct)
            src = x.clone()
            out = src if inplace else torch.empty_like(src)
            R.c2c(src, out, [axis], fwd, fct)
            torch.cuda.synchronize()
            check(out.cpu().numpy(), want, np.float32, n, ("fused four-step", shape, fwd, inplace))
            if not inplace:
                assert torch.equal(src, x)  # the input is left alone
    # output with a different row pitch than the input (the c2r path transforms into a contiguous temporary)
    wide = torch.zeros(n, 900, dtype=torch.complex64, device=dev)
    xh = cplx(rng, (n, 801), np.complex64)
    x = torch.from_numpy(xh).to(dev)
    R.c2c(x, wide[:, 7 : 7 + 801], [0], True, 1.0)
    torch.cuda.synchronize()
    want = np.empty_like(xh)
    T.c2c(xh, want, [0], True, 1.0)
    check(wide[:, 7 : 7 + 801].cpu().numpy(), want, np.float32, n, "fused four-step, different pitches")
    assert float(wide[:, :7].abs().max()) == 0.0 and float(wide[:, 808:].abs().max()) == 0.0


def test_streamed_strided_lines_512_1024(R):
    """512- and 1024-point complex64 lines along a strided axis with enough tiles to fill the device: the persistent kernel
    fed by the copy engine, two warp groups sharing one exchange buffer (pow2_stream_kernel.cuh) -- full and partial tiles,
    one and two batch dims, in place and out of place, both directions, fct != 1, different pitches on the two sides --
    against the reference on the host; RFB200_STREAM=0 (other process) is covered by test_strided_lines_two_per_thread."""
    import torch

    T = trusted()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(21)
    for shape, axis in (((2, 1024, 4800), 1), ((2, 1024, 4808), 1), ((3, 1024, 3200), 1), ((2, 512, 4808), 1), ((5, 512, 1936), 1),
                        ((2, 1024, 3, 1616), 1)):
        n = shape[axis]
        xh = cplx(rng, shape, np.complex64)
        x = torch.from_numpy(xh).to(dev)
        for fwd, fct, inplace in ((True, 1.0, False), (False, 0.5, True)):
            want = np.empty_like(xh)
            T.c2c(xh, want, [axis], fwd, fct)
            src = x.clone()
            out = src if inplace else torch.empty_like(src)
            R.launch_trace(True)
            R.c2c(src, out, [axis], fwd, fct)
            torch.cuda.synchronize()
            names = R.launch_trace_get()
            R.launch_trace(False)
            assert len(names) == 1 and "stream" in names[0], names
            check(out.cpu().numpy(), want, np.float32, n, ("streamed lines", shape, fwd, inplace))
            if not inplace:
                assert torch.equal(src, x)
    # a wider output array (different pitch), untouched outside the written columns
    xh = cplx(rng, (2, 1024, 4800), np.complex64)
    x = torch.from_numpy(xh).to(dev)
    wide = torch.zeros(2, 1024, 4900, dtype=torch.complex64, device=dev)
    R.c2c(x, wide[:, :, 50:4850], [1], True, 1.0)
    torch.cuda.synchronize()
    want = np.empty_like(xh)
    T.c2c(xh, want, [1], True, 1.0)
    check(wide[:, :, 50:4850].cpu().numpy(), want, np.float32, 1024, "streamed lines, different pitches")
    assert float(wide[:, :, :50].abs().max()) == 0.0 and float(wide[:, :, 4850:].abs().max()) == 0.0
    # rows more than 64 KiB apart stay on the register kernel (the copy engine is slow when every row lies on its own page)
    xh = cplx(rng, (1024, 9610), np.complex64)
    x = torch.from_numpy(xh).to(dev)
    R.launch_trace(True)
    R.c2c(x, x, [0], True, 1.0)
    torch.cuda.synchronize()
    names = R.launch_trace_get()
    R.launch_trace(False)
    assert len(names) == 1 and "pair" in names[0], names
    want = np.empty_like(xh)
    T.c2c(xh, want, [0], True, 1.0)
    check(x.cpu().numpy(), want, np.float32, 1024, "far-strided lines")
    # 1024^3-like volume slice: all three axes of (64, 1024, 1024) through the N-D driver
    xh = cplx(rng, (64, 1024, 1024), np.complex64)
    x = torch.from_numpy(xh).to(dev)
    want = np.empty_like(xh)
    T.c2c(xh, want, [0, 1, 2], True, 1.0)
    R.c2c(x, x, [0, 1, 2], True, 1.0)
    torch.cuda.synchronize()
    check(x.cpu().numpy(), want, np.float32, 64 * 1024 * 1024, "volume (64, 1024, 1024)")


def test_r2c_three_axes_goes_through_the_padded_scratch(R):
    """r2c over three (and four) axes whose half-spectrum rows are >= 1 KiB and not a multiple of 128 bytes: the passes between
    the real transform and the last one run on a scratch copy with padded rows (ops.cu r2c_into) -- against the reference,
    float32 and float64, repeated axes included; and the c2r way back (padded temporary)."""
    T = trusted()
    rng = np.random.default_rng(41)
    for dt in (np.float32, np.float64):
        for shape, axes in (((24, 40, 300), [0, 1, 2]), ((6, 20, 24, 258), [1, 2, 3]), ((24, 40, 300), [1, 0, 2]),
                            ((16, 24, 300), [2, 0, 1, 2]), ((3, 8, 12, 1026), [0, 1, 2, 3])):
            x = rng.standard_normal(shape).astype(dt)
            oshape = list(shape)
            oshape[axes[-1]] = shape[axes[-1]] // 2 + 1
            want = np.zeros(oshape, dtype=CD[dt])
            got = np.zeros(oshape, dtype=CD[dt])
            T.r2c(x, want, axes, True, 0.5)
            R.r2c(x, got, axes, True, 0.5)
            n = int(np.prod([shape[a] for a in set(axes)]))
            check(got, want, dt, n, ("r2c padded scratch", shape, axes, dt))
            back_w, back_g = np.empty_like(x), np.empty_like(x)
            T.c2r(want, back_w, axes, False, 1.0)
            R.c2r(want, back_g, axes, False, 1.0)
            check(back_g, back_w, dt, n, ("c2r padded temporary", shape, axes, dt))


def test_strided_lines_two_per_thread(R):
    """float32 lines of 128..1024 points along a strided axis (neighbouring lines adjacent): the kernel that keeps two
    lines per thread (pow2_pair_kernel.cuh) -- odd and even numbers of lines, partial tiles, extra batch dims on either
    side, in place, both directions, fct != 1, and as the two steps of the four-step split (n = 16384, 65536)."""
    T = trusted()
    rng = np.random.default_rng(13)
    cases = []
    for n in (128, 256, 512, 1024):
        cases += [((n, 37), 0), ((n, 64), 0), ((3, n, 45), 1), ((n, 5, 33), 0), ((2, n, 1), 1), ((n, 2), 0)]
    cases += [((16384, 41), 0), ((65536, 40), 0), ((2, 4096, 70), 1)]
    for shape, axis in cases:
        x = cplx(rng, shape, np.complex64)
        n = shape[axis]
        for fwd, fct, inplace in ((True, 1.0, False), (False, 0.25, True)):
            want = np.empty_like(x)
            T.c2c(x, want, [axis], fwd, fct)
            src = x.copy()
            out = src if inplace else np.zeros_like(x)
            R.c2c(src, out, [axis], fwd, fct)
            check(out, want, np.float32, n, ("pair", shape, axis, fwd, inplace))
    # several axes in one call (every pass after the first works in place on the output)
    x = cplx(rng, (128, 256, 48), np.complex64)
    want, got = np.empty_like(x), np.empty_like(x)
    T.c2c(x, want, [0, 1], True, 1.0)
    R.c2c(x, got, [0, 1], True, 1.0)
    check(got, want, np.float32, 128 * 256, "pair, two strided axes")


def test_long_complex_lines_two_transforms_per_thread(R):
    """8192- and 16384-point complex64 contiguous lines (the c2c mode of pow2_dual_kernel.cuh when enabled for the
    length, the single-transform kernel otherwise / for views that are only 8-byte aligned): both directions, in place,
    padded rows, fct != 1 -- against the reference."""
    import torch

    T = trusted()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(14)
    for n in (8192, 16384):
        base_h = cplx(rng, (5, n + 6), np.complex64)
        base = torch.from_numpy(base_h).to(dev)
        for view in (base[:, :n], base[:, 2 : n + 2], base[:, 1 : n + 1], base[3, :n]):
            xh = np.ascontiguousarray(view.cpu().numpy())
            for fwd, fct, inplace in ((True, 1.0, False), (False, 1.0 / n, False), (True, 0.5, True)):
                want = np.empty_like(xh)
                T.c2c(xh, want, [xh.ndim - 1], fwd, fct)
                if inplace:
                    # the same view of a fresh copy of the padded array, transformed in place
                    full = base.clone()
                    off = (view.data_ptr() - base.data_ptr()) // 8
                    src = torch.as_strided(full, view.shape, view.stride(), off)
                    R.c2c(src, src, [src.dim() - 1], fwd, fct)
                    got = src
                else:
                    src = view
                    got = torch.empty(view.shape, dtype=torch.complex64, device=dev)
                    R.c2c(src, got, [src.dim() - 1], fwd, fct)
                torch.cuda.synchronize()
                check(got.cpu().numpy(), want, np.float32, n, ("c2c long", n, tuple(view.stride()), fwd, inplace))
    x = torch.from_numpy(cplx(rng, (2000, 8192), np.complex64)).to(dev)
    y = torch.empty_like(x)
    R.c2c(x, y, [1], True, 1.0)
    xh = x.cpu().numpy()
    want = np.empty_like(xh)
    T.c2c(xh, want, [1], True, 1.0)
    check(y.cpu().numpy(), want, np.float32, 8192, "c2c batch of 2000 lines of 8192")
