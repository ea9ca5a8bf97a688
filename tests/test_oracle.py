"""CPU tests: pin the NumPy oracle (oracle/pocketfft_oracle.py) against
 (a) the committed golden vectors generated from the unmodified reference,
 (b) the compiled reference itself (oracle/_ref) when it is present,
 (c) the O(n^2) definition, the README example and FFTW's DCT/DST known answers."""
import numpy as np
import pytest

import parity
from oracle import pocketfft_oracle as O


def test_golden_cases_match_oracle():
    d, cases = parity.golden()
    worst = 0.0
    for i, c in enumerate(cases):
        ain = d[f"c{i}_in"]
        want = d[f"c{i}_out"]
        got = np.zeros_like(want)
        parity.call(O, c, ain, got)
        single = want.dtype in (np.float32, np.complex64)
        # golden = reference computed in its own precision; oracle computes in fp64
        bound = 3e-6 if single else 1e-13
        e = parity.l2err(got, want)
        worst = max(worst, e)
        assert e < bound, (i, c, e)
    assert worst > 0.0


def test_good_size_golden():
    d, _ = parity.golden()
    for t, gc, gr in zip(d["good_size_targets"], d["good_size_cmplx"], d["good_size_real"]):
        assert O.good_size(int(t), False) == int(gc)
        assert O.good_size(int(t), True) == int(gr)


def test_readme_example():
    d, _ = parity.golden()
    out = np.empty(8, dtype=np.complex128)
    O.c2c(d["readme_in"], out, [0], True, 1.0)
    assert parity.l2err(out, d["readme_out"]) < 1e-15
    assert abs(out[0] - 31) < 1e-12 and abs(out[2] - (-3 + 8j)) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 12, 17, 30, 49, 64, 127, 210])
def test_engine_vs_definition(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for fwd in (True, False):
        out = np.empty_like(x)
        O.c2c(x, out, [0], fwd, 1.0)
        assert parity.l2err(out, O.dft_direct(x, fwd)) < 5e-15


def test_fftw_dct_dst_known_answers():
    d, _ = parity.golden()
    keys = [k for k in d.files if k.startswith("fftw_")]
    assert len(keys) >= 60
    for k in keys:
        _, kind, t, N = k.split("_")
        t, N = int(t), int(N)
        x = np.linspace(0, N - 1, N)
        y = np.empty_like(x)
        getattr(O, kind)(x, y, [0], t, 1.0, False)
        want = d[k]
        # FFTW's REDFT/RODFT share SciPy's unnormalised convention except for a
        # factor the SciPy tests divide out for type 1 (none needed here: check scale)
        e = parity.l2err(y, want)
        assert e < 1e-12, (k, e)


needs_ref = pytest.mark.skipif(parity.reflib() is None, reason="oracle/_ref not built")


@needs_ref
def test_good_size_sweep_vs_reference():
    ref = parity.reflib()
    for n in list(range(0, 20000)) + [15015, 1000003, 2000005, 2**31 - 1, 2**31 + 1, 2**40 + 1]:
        for r in (False, True):
            assert O.good_size(n, r) == ref.good_size(n, r)


@needs_ref
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_oracle_vs_reference_sweep(dt):
    ref = parity.reflib()
    rng = np.random.default_rng(7)
    cdt = np.complex128 if dt is np.float64 else np.complex64
    bound = 1e-13 if dt is np.float64 else 3e-6
    for n in list(range(1, 130)) + [255, 256, 257, 1000, 1021, 2048]:
        x = rng.standard_normal(n).astype(dt)
        z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(cdt)
        for fwd in (True, False):
            a, b = np.empty_like(z), np.empty_like(z)
            ref.c2c(z, a, [0], fwd, 1.0)
            O.c2c(z, b, [0], fwd, 1.0)
            assert parity.l2err(b, a) < bound
            a = np.zeros(n // 2 + 1, dtype=cdt)
            b = np.zeros_like(a)
            ref.r2c(x, a, [0], fwd, 1.0)
            O.r2c(x, b, [0], fwd, 1.0)
            assert parity.l2err(b, a) < bound
            zz = z[: n // 2 + 1].copy()
            a, b = np.empty(n, dtype=dt), np.empty(n, dtype=dt)
            ref.c2r(zz, a, [0], fwd, 1.0)
            O.c2r(zz, b, [0], fwd, 1.0)
            assert parity.l2err(b, a) < bound
            for r2h in (True, False):
                a, b = np.empty_like(x), np.empty_like(x)
                ref.r2r_fftpack(x, a, [0], r2h, fwd, 1.0)
                O.r2r_fftpack(x, b, [0], r2h, fwd, 1.0)
                assert parity.l2err(b, a) < bound, (n, r2h, fwd)
        a, b = np.empty_like(x), np.empty_like(x)
        ref.r2r_separable_hartley(x, a, [0], 1.0)
        O.r2r_separable_hartley(x, b, [0], 1.0)
        assert parity.l2err(b, a) < bound
        if n > 600:
            continue
        for kind in ("dct", "dst"):
            for t in (1, 2, 3, 4):
                if kind == "dct" and t == 1 and n < 2:
                    continue
                for ortho in (False, True):
                    a, b = np.empty_like(x), np.empty_like(x)
                    getattr(ref, kind)(x, a, [0], t, 1.0, ortho)
                    getattr(O, kind)(x, b, [0], t, 1.0, ortho)
                    assert parity.l2err(b, a) < bound * 5, (kind, t, n, ortho)


@needs_ref
def test_oracle_vs_reference_nd_and_strides():
    ref = parity.reflib()
    rng = np.random.default_rng(11)
    base = rng.standard_normal((12, 10, 9)) + 1j * rng.standard_normal((12, 10, 9))
    views = [base, np.asfortranarray(base), base[::-1, :, ::-1], base[::2, ::3, :]]
    for v in views:
        for axes in ([0], [1], [2], [0, 1, 2], [2, 1, 0], [1, 1], [2, 0], [0, 0, 1]):
            a = np.empty(v.shape, dtype=np.complex128)
            b = np.empty(v.shape, dtype=np.complex128)
            ref.c2c(v, a, axes, True, 0.5)
            O.c2c(v, b, axes, True, 0.5)
            assert parity.l2err(b, a) < 1e-13
            x = np.ascontiguousarray(v.real)
            a = np.empty_like(x)
            b = np.empty_like(x)
            ref.r2r_separable_hartley(x, a, axes, 1.0)
            O.r2r_separable_hartley(x, b, axes, 1.0)
            assert parity.l2err(b, a) < 1e-13, axes
            oshp = list(x.shape)
            oshp[axes[-1]] = oshp[axes[-1]] // 2 + 1
            ac = np.zeros(oshp, dtype=np.complex128)
            bc = np.zeros(oshp, dtype=np.complex128)
            ref.r2c(x, ac, axes, False, 1.0)
            O.r2c(x, bc, axes, False, 1.0)
            assert parity.l2err(bc, ac) < 1e-13, axes
            zin = ac.copy()
            ref.c2r(zin, a, axes, False, 1.0)
            O.c2r(zin, b, axes, False, 1.0)
            assert parity.l2err(b, a) < 1e-13, axes
            if len(set(axes)) != len(axes):
                continue  # repeated axes: the mirror-based ops depend on iteration order (unspecified)
            ref.r2r_genuine_hartley(x, a, axes, 1.0)
            O.r2r_genuine_hartley(x, b, axes, 1.0)
            assert parity.l2err(b, a) < 1e-13, axes
            ac = np.empty(x.shape, dtype=np.complex128)
            bc = np.empty(x.shape, dtype=np.complex128)
            ref.c2c_sym(x, ac, axes, True, 1.0)
            O.c2c_sym(x, bc, axes, True, 1.0)
            assert parity.l2err(bc, ac) < 1e-13, axes
