"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of what each ``numba_*`` entry point of the reference computes
(reference boundary: rocket_fft/_pocketfft_numba.cpp:25-223; engine:
rocket_fft/_pocketfft_hdronly.h).  Every function states the transform as
mathematics (a sum over the input) and evaluates it in float64/complex128 whatever
the I/O dtype, then casts to the output dtype.  The 1-D complex DFT it is built on
is ``numpy.fft`` (NumPy >= 2 ships its own C++ PocketFFT, bit-identical to the
reference for c2c) and, for small lengths, ``dft_direct`` -- the O(n^2) definition in
extended precision -- which pins the fast path in tests/test_oracle.py.

PARITY PINNING: tests/test_oracle.py checks this module (a) against the compiled,
unmodified reference in oracle/_ref (built by oracle/Makefile from /root/reference),
(b) against the committed golden vectors in tests/golden/ that were generated from
that reference by tests/golden/make_golden.py, and (c) against the reference's own
known-answer material (README.md:26-35 example; FFTW DCT/DST vectors used by
tests/test_scipy_testsuite.py:1192-1293).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference
legs may import this package.  The product (rocket_fft_b200/) never does.
"""
from __future__ import annotations

import numpy as np

_SQRT2 = np.sqrt(2.0)


# ----------------------------------------------------------------------------
# good_size   (reference: _pocketfft_hdronly.h:589-650, boundary B:25-29)
# ----------------------------------------------------------------------------
def good_size(target: int, real: bool) -> int:
    """Smallest integer >= target whose prime factors are all in {2,3,5,7,11}
    (complex) or {2,3,5} (real).  Targets <= 12 (complex) / <= 6 (real) are returned
    unchanged (H:592, 625), including 0."""
    target = int(target)
    primes = (2, 3, 5) if real else (2, 3, 5, 7, 11)
    if target <= (6 if real else 12):
        return target
    best = 1
    while best < target:  # a power of two always qualifies
        best *= 2

    def rec(idx, val):
        nonlocal best
        if val >= target:
            best = min(best, val)
            return
        if idx == len(primes):
            return
        v = val
        while v < best:
            rec(idx + 1, v)
            v *= primes[idx]

    rec(0, 1)
    return int(best)


# ----------------------------------------------------------------------------
# 1-D complex engine
# ----------------------------------------------------------------------------
def dft_direct(x, forward=True, axis=-1):
    """X[k] = sum_j x[j] exp(-/+ 2 pi i j k / n) evaluated literally in long double."""
    x = np.moveaxis(np.asarray(x), axis, -1)
    n = x.shape[-1]
    j = np.arange(n)
    jk = np.outer(j, j) % n
    ang = (2 * np.pi * jk.astype(np.longdouble)) / np.longdouble(n)
    s = -1 if forward else 1
    wr = np.cos(ang)
    wi = s * np.sin(ang)
    xr = x.real.astype(np.longdouble)
    xi = x.imag.astype(np.longdouble) if np.iscomplexobj(x) else np.zeros_like(xr)
    yr = xr @ wr - xi @ wi
    yi = xr @ wi + xi @ wr
    y = yr.astype(np.float64) + 1j * yi.astype(np.float64)
    return np.moveaxis(y, -1, axis)


def _cfft(x, axis, forward):
    """Unnormalised complex DFT along one axis, sign -1 if forward else +1."""
    x = np.asarray(x, dtype=np.complex128)
    n = x.shape[axis]
    if forward:
        return np.fft.fft(x, axis=axis)
    return np.fft.ifft(x, axis=axis) * n


def _out(arr, aout):
    """Write into the caller's output array (any strides), casting to its dtype."""
    if np.iscomplexobj(aout):
        aout[...] = arr.astype(aout.dtype)
    else:
        aout[...] = np.real(arr).astype(aout.dtype)
    return aout


def _any_zero(shape):
    return any(s == 0 for s in shape)


# ----------------------------------------------------------------------------
# c2c   (B:31-49 -> H:3875-3889 -> general_nd H:3568-3607)
# ----------------------------------------------------------------------------
def c2c(ain, aout, axes, forward, fct, nthreads=1):
    """For each axis in ``axes`` in the given order (repeats allowed):
    Y[k] = sum_j X[j] exp(s 2 pi i j k / n), s=-1 if forward else +1; result * fct."""
    if _any_zero(ain.shape):
        return aout
    y = np.asarray(ain, dtype=np.complex128)
    for ax in axes:
        y = _cfft(y, int(ax), forward)
    return _out(y * fct, aout)


# ----------------------------------------------------------------------------
# r2c   (B:91-109 -> H:3955-3975; general_r2c H:3723-3779)
# ----------------------------------------------------------------------------
def _r2c_core(ain, axes, forward, fct):
    x = np.asarray(ain, dtype=np.float64)
    last = int(axes[-1])
    n = x.shape[last]
    y = _cfft(x, last, forward)
    y = np.take(y, np.arange(n // 2 + 1), axis=last) * fct
    for ax in list(axes)[:-1]:
        y = _cfft(y, int(ax), forward)
    return y


def r2c(ain, aout, axes, forward, fct, nthreads=1):
    """Half spectrum (k = 0..n/2) of the real DFT along axes[-1], then full complex
    DFTs (same sign, no extra scaling) along axes[:-1] in order."""
    if _any_zero(ain.shape):
        return aout
    y = _r2c_core(ain, axes, forward, fct)
    idx = tuple(slice(0, s) for s in y.shape)
    aout[idx] = y.astype(aout.dtype)
    return aout


# ----------------------------------------------------------------------------
# c2r   (B:145-163 -> H:3995-4023; general_c2r H:3781-3852)
# ----------------------------------------------------------------------------
def c2r(ain, aout, axes, forward, fct, nthreads=1):
    """Complex DFTs along axes[:-1], then along axes[-1] the real signal whose
    Hermitian half-spectrum is the input:  y[j] = sum_k Xh[k] exp(s 2 pi i j k / n)
    with Xh the Hermitian extension to length n = aout.shape[axis].  The imaginary
    part of bin 0 (and of the Nyquist bin for even n) is ignored (H:3830,3845)."""
    if _any_zero(aout.shape):
        return aout
    last = int(axes[-1])
    n = aout.shape[last]
    nh = n // 2 + 1
    x = np.asarray(ain, dtype=np.complex128)
    x = np.take(x, np.arange(nh), axis=last)
    for ax in list(axes)[:-1]:
        x = _cfft(x, int(ax), forward)
    x = np.moveaxis(x, last, -1)
    full = np.empty(x.shape[:-1] + (n,), dtype=np.complex128)
    full[..., :nh] = x
    full[..., 0] = x[..., 0].real
    if n % 2 == 0:
        full[..., n // 2] = x[..., n // 2].real
    k = np.arange(nh, n)
    full[..., nh:] = np.conj(full[..., n - k])
    y = _cfft(full, -1, forward).real * fct
    y = np.moveaxis(y, -1, last)
    return _out(y, aout)


# ----------------------------------------------------------------------------
# c2c_sym   (B:111-143: r2c + conjugate mirror over all transformed axes)
# ----------------------------------------------------------------------------
def _mirror_full(half, shape, axes):
    """Expand a half spectrum (k_L = 0..n_L/2 along L = axes[-1]) to the full array:
    entries with k_L > n_L/2 are conj(half[-k]) with the index negated (mod n) along
    every axis in set(axes)  (rev_iter, H:3383-3444).  The reference walks the half
    space and writes each value's mirror; where both an index and its mirror lie in
    the half space (k_L in {0, n_L/2}) the two agree for Hermitian-consistent data,
    which is every case except *repeated axes* -- there the reference's result
    depends on its iteration order and is left unspecified here."""
    L = int(axes[-1])
    n = shape[L]
    nh = n // 2 + 1
    full = np.empty(shape, dtype=np.complex128)
    sl = [slice(None)] * len(shape)
    sl[L] = slice(0, nh)
    full[tuple(sl)] = half
    if n > nh:
        # index arrays of the mirrored positions
        idx = []
        for d, s in enumerate(shape):
            r = np.arange(nh, n) if d == L else np.arange(s)
            if d in set(int(a) for a in axes):
                r = (-r) % s
            idx.append(r)
        src = half[np.ix_(*idx)]
        sl[L] = slice(nh, n)
        full[tuple(sl)] = np.conj(src)
    return full


def c2c_sym(ain, aout, axes, forward, fct, nthreads=1):
    """Full complex DFT over ``axes`` of a real input: r2c, then fill the other half
    with the conjugate mirror."""
    if _any_zero(ain.shape):
        return aout
    half = _r2c_core(ain, axes, forward, fct)
    return _out(_mirror_full(half, ain.shape, axes), aout)


# ----------------------------------------------------------------------------
# halfcomplex helpers (FFTPACK layout  [r0, r1, i1, r2, i2, ..., (r_{n/2})])
# ----------------------------------------------------------------------------
def _r2hc(x):
    """real (..., n) -> halfcomplex (..., n) of the forward (sign -) DFT. (H:2573-2627)"""
    n = x.shape[-1]
    X = np.fft.fft(np.asarray(x, dtype=np.float64), axis=-1)
    out = np.empty(x.shape, dtype=np.float64)
    out[..., 0] = X[..., 0].real
    for k in range(1, (n + 1) // 2):
        out[..., 2 * k - 1] = X[..., k].real
        out[..., 2 * k] = X[..., k].imag
    if n % 2 == 0 and n > 0:
        out[..., n - 1] = X[..., n // 2].real
    return out


def _hc2r(h):
    """halfcomplex (..., n) -> real (..., n), backward (sign +), unnormalised."""
    n = h.shape[-1]
    h = np.asarray(h, dtype=np.float64)
    X = np.zeros(h.shape[:-1] + (n,), dtype=np.complex128)
    X[..., 0] = h[..., 0]
    for k in range(1, (n + 1) // 2):
        X[..., k] = h[..., 2 * k - 1] + 1j * h[..., 2 * k]
        X[..., n - k] = h[..., 2 * k - 1] - 1j * h[..., 2 * k]
    if n % 2 == 0 and n > 0:
        X[..., n // 2] = h[..., n - 1]
    return (np.fft.ifft(X, axis=-1) * n).real


def _negimag(h):
    h = h.copy()
    h[..., 2::2] = -h[..., 2::2]
    return h


def _per_axis(ain, axes, fn):
    y = np.asarray(ain, dtype=np.float64)
    for ax in axes:
        ax = int(ax)
        y = np.moveaxis(fn(np.moveaxis(y, ax, -1)), -1, ax)
    return y


# ----------------------------------------------------------------------------
# r2r_fftpack   (B:165-183 -> H:4025-4040, ExecR2R H:3854-3873)
# ----------------------------------------------------------------------------
def r2r_fftpack(ain, aout, axes, real2hermitian, forward, fct, nthreads=1):
    """Per axis.  NOTE the reference's ExecR2R runs the real plan in direction
    ``forward`` (``plan.exec(buf, fct, forward)``, H:3867), not in direction
    ``real2hermitian`` as upstream PocketFFT does, so the four flag combinations give
      (r2h=T, fwd=T)  real -> halfcomplex of sum x e^{-2 pi i jk/n}
      (r2h=F, fwd=F)  halfcomplex -> real with e^{+2 pi i jk/n}
      (r2h=T, fwd=F)  hc2r(x) with entries 2,4,6.. negated afterwards
      (r2h=F, fwd=T)  r2hc(x with entries 2,4,6.. negated)
    This oracle restates exactly that (verified against oracle/_ref)."""
    if _any_zero(ain.shape):
        return aout
    r2h = bool(real2hermitian)
    fwd = bool(forward)

    def one(v):
        if (not r2h) and fwd:
            v = _negimag(v)
        v = _r2hc(v) if fwd else _hc2r(v)
        if r2h and (not fwd):
            v = _negimag(v)
        return v

    return _out(_per_axis(ain, axes, one) * fct, aout)


# ----------------------------------------------------------------------------
# Hartley   (B:185-223 -> H:4042-4091; copy_hartley H:3661-3692)
# ----------------------------------------------------------------------------
def r2r_separable_hartley(ain, aout, axes, fct, nthreads=1):
    """Per axis: H[k] = Re F[k] + Im F[k], F the forward DFT along that axis."""
    if _any_zero(ain.shape):
        return aout

    def one(v):
        F = np.fft.fft(v, axis=-1)
        return F.real + F.imag

    return _out(_per_axis(ain, axes, one) * fct, aout)


def r2r_genuine_hartley(ain, aout, axes, fct, nthreads=1):
    """Re F + Im F of the N-D forward DFT over ``axes`` (one axis -> separable)."""
    if _any_zero(ain.shape):
        return aout
    if len(axes) == 1:
        return r2r_separable_hartley(ain, aout, axes, fct, nthreads)
    y = _mirror_full(_r2c_core(ain, axes, True, fct), ain.shape, axes)
    return _out(y.real + y.imag, aout)


# ----------------------------------------------------------------------------
# DCT / DST types 1-4   (B:51-89 -> H:3891-3935; T_dct1 2918, T_dst1 2957,
# T_dcst23 2987, T_dcst4 3063)
# ----------------------------------------------------------------------------
def _dct_def(x, type):
    """Unnormalised definitions along the last axis (SciPy conventions)."""
    N = x.shape[-1]
    n = np.arange(N, dtype=np.float64)
    k = n[:, None]
    if type == 1:
        if N == 1:
            return x.copy()
        M = 2 * np.cos(np.pi * k * n / (N - 1))
        M[:, 0] = 1.0
        M[:, -1] = (-1.0) ** n
    elif type == 2:
        M = 2 * np.cos(np.pi * k * (2 * n + 1) / (2 * N))
    elif type == 3:
        M = 2 * np.cos(np.pi * n * (2 * k + 1) / (2 * N))
        M[:, 0] = 1.0
    elif type == 4:
        M = 2 * np.cos(np.pi * (2 * k + 1) * (2 * n + 1) / (4 * N))
    else:
        raise ValueError("invalid DCT type")
    return x @ M.T


def _dst_def(x, type):
    N = x.shape[-1]
    n = np.arange(N, dtype=np.float64)
    k = n[:, None]
    if type == 1:
        M = 2 * np.sin(np.pi * (n + 1) * (k + 1) / (N + 1))
    elif type == 2:
        M = 2 * np.sin(np.pi * (k + 1) * (2 * n + 1) / (2 * N))
    elif type == 3:
        M = 2 * np.sin(np.pi * (2 * k + 1) * (n + 1) / (2 * N))
        M[:, -1] = (-1.0) ** n
    elif type == 4:
        M = 2 * np.sin(np.pi * (2 * k + 1) * (2 * n + 1) / (4 * N))
    else:
        raise ValueError("invalid DST type")
    return x @ M.T


def _dcst_fast(x, type, cosine):
    """Same transforms through scipy.fft when available (O(n log n)) -- used for
    long lines; falls back to the O(n^2) definition."""
    try:
        import scipy.fft as sf
    except Exception:  # pragma: no cover
        return _dct_def(x, type) if cosine else _dst_def(x, type)
    if x.shape[-1] <= 64:
        return _dct_def(x, type) if cosine else _dst_def(x, type)
    f = sf.dct if cosine else sf.dst
    return f(x, type=type, axis=-1, norm=None)


def _dcst(ain, aout, axes, type, fct, ortho, cosine):
    if _any_zero(ain.shape):
        return aout
    type = int(type)
    if type not in (1, 2, 3, 4):
        raise ValueError("invalid DCT/DST type")

    def one(v):
        v = np.array(v, dtype=np.float64)
        N = v.shape[-1]
        if ortho:
            # H:2934-2937 (DCT-I), H:3038-3039 (type 3: element 0, for DST too -- the
            # reference's documented quirk, README.md:61-65)
            if type == 1 and cosine:
                v[..., 0] *= _SQRT2
                v[..., N - 1] *= _SQRT2
            elif type == 3:
                v[..., 0] *= _SQRT2
        y = _dcst_fast(v, type, cosine)
        if ortho:
            if type == 1 and cosine:
                y[..., 0] *= _SQRT2 * 0.5
                y[..., N - 1] *= _SQRT2 * 0.5
            elif type == 2:
                y[..., 0] *= _SQRT2 * 0.5  # H:3033-3034 (element 0 for DST as well)
        return y

    # fct is applied inside the real plan on every axis?  No: general_nd passes fct to
    # the first axis only (H:3605) -- a single overall factor.
    return _out(_per_axis(ain, axes, one) * fct, aout)


def dct(ain, aout, axes, type, fct, ortho, nthreads=1):
    return _dcst(ain, aout, axes, type, fct, ortho, True)


def dst(ain, aout, axes, type, fct, ortho, nthreads=1):
    return _dcst(ain, aout, axes, type, fct, ortho, False)
