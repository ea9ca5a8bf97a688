"""Robustness of the library around the transform path (GPU box): bounded plan cache with eviction, behaviour of a
forked child (reference: the pthread_atfork handlers of its thread pool, _pocketfft_hdronly.h:992-1006, exercised by
the reference's tests/test_multiprocessing.py:27-35), arrays on a non-current device, the failure channel of the void
numba_* entry points."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


def _run(code, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, timeout=timeout, cwd=parity.ROOT, env=e)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout


def test_plan_cache_is_bounded_and_evicts_least_recently_used():
    out = _run(
        """
        import numpy as np, rocket_fft_b200 as R
        from oracle import pocketfft_oracle as O
        rng = np.random.default_rng(0)
        worst = 0.0
        lengths = [64, 96, 100, 128, 243, 256, 360, 500, 512, 1000, 1024, 1009, 2048, 3000, 4096, 4099, 5000, 8192, 15015, 16384]
        for rep in range(2):
            for n in lengths:
                x = (rng.standard_normal((4, n)) + 1j * rng.standard_normal((4, n))).astype(np.complex64)
                a, b = np.empty_like(x), np.empty_like(x)
                R.c2c(x, a, [1], True, 1.0)
                O.c2c(x, b, [1], True, 1.0)
                worst = max(worst, float(np.linalg.norm(a - b) / np.linalg.norm(b)))
                e, nbytes = R.plan_cache_stats()
                assert e <= 6 + 4, (n, e)   # the bound, plus the tables leased by the call that just ran
        print("worst", worst, "entries", R.plan_cache_stats())
        assert worst < 2e-4
        R.plan_cache_clear()
        assert R.plan_cache_stats()[0] == 0
        """,
        env={"RFB200_PLAN_CACHE_ENTRIES": "6"},
    )
    assert "worst" in out


def test_forked_child_gets_a_clear_error_not_undefined_behaviour():
    out = _run(
        """
        import os, sys, numpy as np, rocket_fft_b200 as R
        x = np.ones((4, 64), dtype=np.complex64)
        y = np.empty_like(x)
        R.c2c(x, y, [1], True, 1.0)            # the parent touches CUDA
        assert abs(y[0, 0] - 64) < 1e-3
        r, w = os.pipe()
        pid = os.fork()
        if pid == 0:
            msg = "no error"
            try:
                z = np.zeros_like(x)
                R.c2c(x, z, [1], True, 1.0)
            except R.TransformError as e:
                msg = "TransformError: " + str(e) + (" nan" if np.isnan(z.real).all() else " not-nan")
            except BaseException as e:
                msg = "other: " + repr(e)
            os.write(w, msg.encode())
            os._exit(0)
        os.close(w)
        _, status = os.waitpid(pid, 0)
        print("child status", status, "|", os.read(r, 4096).decode())
        R.c2c(x, y, [1], True, 1.0)            # the parent is unaffected
        assert abs(y[0, 0] - 64) < 1e-3
        """
    )
    assert "child status 0" in out and "TransformError" in out and "fork" in out and " nan" in out, out


def test_a_child_forked_before_the_first_transform_works():
    out = _run(
        """
        import os, numpy as np, rocket_fft_b200 as R
        assert R.good_size(1000003, False) == 1000188   # host-only: does not touch CUDA
        x = np.ones((4, 64), dtype=np.complex64)
        pid = os.fork()
        if pid == 0:
            y = np.empty_like(x)
            R.c2c(x, y, [1], True, 1.0)
            os._exit(0 if abs(y[0, 0] - 64) < 1e-3 else 3)
        _, status = os.waitpid(pid, 0)
        print("child status", status)
        """
    )
    assert "child status 0" in out, out


def test_arrays_on_a_device_that_is_not_current():
    import torch

    import rocket_fft_b200 as R

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(5)
    xh = (rng.standard_normal((64, 1024)) + 1j * rng.standard_normal((64, 1024))).astype(np.complex64)
    want = np.empty_like(xh)
    (parity.reflib() or __import__("oracle.pocketfft_oracle", fromlist=["x"])).c2c(xh, want, [1], True, 1.0)
    x = torch.from_numpy(xh).to("cuda:1")
    y = torch.empty_like(x)
    with torch.cuda.device(0):
        R.c2c(x, y, [1], True, 1.0)  # cuda:0 is current, the arrays live on cuda:1
        torch.cuda.synchronize(1)
        assert torch.cuda.current_device() == 0
    assert parity.l2err(y.cpu().numpy(), want) < parity.tol(np.float32, 1024)
    with pytest.raises((ValueError, R.TransformError)):
        R.c2c(x, torch.empty_like(x, device="cuda:0"), [1], True, 1.0)


def test_failed_void_call_fills_nan_and_counts():
    import rocket_fft_b200 as R

    before = R.failure_count()
    x = np.ones((3, 8), dtype=np.float64)
    y = np.zeros((3, 8), dtype=np.float64)
    with pytest.raises(R.TransformError):
        R.lib.dct(x[:, :1], y[:, :1], [1], 1, 1.0, False)  # DCT-I of one point: an error in the reference as well
        err = R.last_error()
        if err:
            raise R.TransformError(err)
    assert R.failure_count() == before + 1
    assert np.isnan(y[:, 0]).all() and (y[:, 1:] == 0).all()  # only the output's own elements are touched


def test_dst_ortho_scaling_per_call():
    """numba_dst keeps the reference's DST-II/III ortho scaling (element 0); dst_ortho="scipy" gives SciPy's for one call."""
    import scipy.fft

    import rocket_fft_b200 as R

    rng = np.random.default_rng(6)
    x = rng.standard_normal((5, 24))
    ref = parity.reflib()
    for t in (2, 3):
        n = x.shape[1]
        fct = 1.0 / np.sqrt(2.0 * n)
        got, want = np.empty_like(x), np.empty_like(x)
        R.dst(x, got, [1], t, fct, True)
        if ref is not None:
            ref.dst(x, want, [1], t, fct, True, 1)
            assert parity.l2err(got, want) < 1e-12
        R.dst(x, got, [1], t, fct, True, dst_ortho="scipy")
        assert parity.l2err(got, scipy.fft.dst(x, t, axis=1, norm="ortho")) < 1e-12
        import torch

        d = torch.from_numpy(x).cuda()
        o = torch.empty_like(d)
        R.dst(d, o, [1], t, fct, True, dst_ortho="scipy")
        assert parity.l2err(o.cpu().numpy(), scipy.fft.dst(x, t, axis=1, norm="ortho")) < 1e-12
    from rocket_fft_b200 import fft as F

    for t in (1, 2, 3, 4):
        assert parity.l2err(F.dst(x, t, norm="ortho"), scipy.fft.dst(x, t, norm="ortho")) < 1e-12
        assert parity.l2err(F.idst(x, t, norm="ortho"), scipy.fft.idst(x, t, norm="ortho")) < 1e-12
        assert parity.l2err(F.dstn(x, t, norm="ortho"), scipy.fft.dstn(x, t, norm="ortho")) < 1e-12
    assert np.allclose(F.fft([1.0, 2.0, 3.0]), np.fft.fft([1.0, 2.0, 3.0]))  # array-likes are accepted


def test_caller_pinned_numpy_arrays():
    """rfb200_host_pin / rocket_fft_b200.pinned: NumPy arrays page-locked by the caller go through the host path without the
    staging ring and give the same result; pinning the same range twice is an error that leaves the first pin intact."""
    out = _run(
        """
        import numpy as np, rocket_fft_b200 as R
        from oracle import pocketfft_oracle as O
        rng = np.random.default_rng(3)
        x = rng.standard_normal((8, 1 << 20)).astype(np.float32)     # 32 MiB: above the staging threshold
        got = np.empty((8, (1 << 19) + 1), dtype=np.complex64)
        want = np.empty_like(got)
        O.r2c(x, want, [1], True, 1.0)
        with R.pinned(x, got):
            R.r2c(x, got, [1], True, 1.0)
            err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            try:
                with R.pinned(x):
                    print("second pin succeeded")
            except R.TransformError as e:
                print("second pin refused:", str(e)[:60])
            got[...] = 0
            R.r2c(x, got, [1], True, 1.0)
            err = max(err, float(np.linalg.norm(got - want) / np.linalg.norm(want)))
        got[...] = 0
        R.r2c(x, got, [1], True, 1.0)                                 # pageable again: the staging ring
        err = max(err, float(np.linalg.norm(got - want) / np.linalg.norm(want)))
        print("err", err)
        assert err < 1e-5 * 20
        """
    )
    assert "err" in out and "refused" in out, out
