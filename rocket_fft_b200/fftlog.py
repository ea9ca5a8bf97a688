"""Fast Hankel transform on logarithmically spaced samples (FFTLog): scipy.fft.fht / ifht / fhtoffset
on top of the transform path -- a caller of r2c / c2r (reference:
rocket_fft/overloads.py:858-902 `fhtcoeff`, `_fhtq`, `_ifhtq` and 1755-1859 the scipy.fft overloads).

    A = fht(a):   A_k = sum_m  u_m  a^_m  exp(...)      with  a^ = rfft(a), then the result is read backwards

How it runs here, per call (bias = 0): three kernel launches and no copies --
  1. r2c along the last axis                                         (transform path)
  2. spectrum *= u'   with u'_m = u_m exp(-2 pi i m / n)             (rfb200_scale_lines)
  3. c2r with forward=True and fct = 1/n                             (transform path)
The reference computes irfft(A u) and returns it reversed along the last axis (`A[..., ::-1]`).  Reversal
z[j] = x[n-1-j] = x[-(j+1)] is conjugation plus a one-sample shift in the frequency domain:
z = irfft(conj(A u) e^{+2 pi i m/n}) = c2r_forward(A u e^{-2 pi i m/n}) / n, because c2r with forward=True
is the transform of the conjugated half-spectrum (the hfft convention, _pocketfft_hdronly.h:3995-4023).
So the phase goes into the coefficient table and the reversed copy disappears.

The coefficients u_m (m = 0..n/2) are plan data like twiddles: O(n) values computed once on the host
in double precision with scipy.special.loggamma / poch -- the very functions the reference binds
(rocket_fft/_special_helpers.cpp:67-69 imports them from scipy.special.cython_special).
"""
from __future__ import annotations

import math
import sys

import numpy as np

from . import lowlevel as _ll

_LN2 = math.log(2.0)


def fhtcoeff(n, dln, mu, offset=0.0, bias=0.0):
    """Coefficient vector u_m, m = 0..n//2, of the fast Hankel transform (complex128, host).

        ln u_m = [Re lnG(x+ + i y_m) - Re lnG(x- + i y_m) + q ln 2]
                 + i [Im lnG(x+ + i y_m) + Im lnG(x- + i y_m) + 2 y_m (ln 2 - offset)]

    with q = bias, x+- = (mu + 1 +- q)/2, y_m = pi m / (n dln), lnG = log-gamma; the imaginary part of the
    last coefficient is cleared for every n like the reference does (SciPy >= 1.12 only for even n), and a
    non-finite u_0 (a pole of the gamma ratio) is replaced by 2^q poch(x-, x+ - x-) -- which may itself be
    infinite: the singular transform (reference: O:859-877)."""
    from scipy.special import loggamma, poch

    n = int(n)
    q = float(bias)
    xp = (mu + 1.0 + q) / 2.0
    xm = (mu + 1.0 - q) / 2.0
    m = n // 2
    y = np.linspace(0.0, np.pi * m / (n * dln), m + 1)
    lg_m = loggamma(xm + 1j * y)
    lg_p = loggamma(xp + 1j * y)
    phase = lg_p.imag + lg_m.imag + y * (2.0 * (_LN2 - offset))
    with np.errstate(all="ignore"):
        u = np.exp((lg_p.real - lg_m.real + _LN2 * q) + 1j * phase)
    u.imag[-1] = 0.0
    if not np.isfinite(u[0]):
        u[0] = 2.0 ** q * poch(xm, xp - xm)
    return u


def fhtoffset(dln, mu, initial=0.0, bias=0.0):
    """Offset close to `initial` that satisfies the low-ringing condition (scipy.fft.fhtoffset; O:1845-1859)."""
    from scipy.special import loggamma

    q = float(bias)
    xp = (mu + 1.0 + q) / 2.0
    xm = (mu + 1.0 - q) / 2.0
    y = np.pi / (2.0 * dln)
    zp = loggamma(xp + 1j * y)
    zm = loggamma(xm + 1j * y)
    arg = (_LN2 - initial) / dln + (zp.imag + zm.imag) / np.pi
    return float(initial + (arg - np.round(arg)) * dln)


def _real_dtype(dt):
    dt = np.dtype(dt)
    if dt.kind == "c":
        raise TypeError("fht / ifht work on real arrays")
    return np.dtype(np.float32) if dt in (np.dtype(np.float32), np.dtype(np.float16)) else np.dtype(np.float64)


def _to_device(a):
    import torch

    if isinstance(a, np.ndarray) or not type(a).__module__.startswith("torch"):
        a = np.asarray(a)
        rdt = _real_dtype(a.dtype)
        return torch.from_numpy(np.ascontiguousarray(a, dtype=rdt)).cuda(), True, rdt
    rdt = _real_dtype(str(a.dtype).replace("torch.", ""))
    return a.to(getattr(torch, rdt.name)), False, rdt


def _hankel(a, dln, mu, offset, bias, inverse):
    import torch

    x, host, rdt = _to_device(a)
    if x.dim() == 0:
        raise ValueError("fht / ifht need at least one dimension")
    # like the reference, the scalar arguments take the precision of the data (O:1764-1768)
    dln, mu, offset, bias = (float(rdt.type(v)) for v in (dln, mu, offset, bias))
    n = x.shape[-1]
    cdt = torch.complex64 if rdt == np.dtype(np.float32) else torch.complex128
    tdt = getattr(torch, rdt.name)
    if n == 0 or x.numel() == 0:
        out = torch.empty_like(x)
        return out.cpu().numpy() if host else out
    x = x.contiguous()
    work = x
    owned = False
    jc = (n - 1) / 2.0
    j = np.arange(n, dtype=np.float64)
    if bias != 0.0:
        pre = np.exp(bias * ((j - jc) * dln + offset)) if inverse else np.exp(-bias * (j - jc) * dln)
        work = x.clone()
        owned = True
        _ll.scale_lines(work, torch.from_numpy(pre.astype(rdt)).to(x.device))
    u = fhtcoeff(n, dln, mu, offset=offset, bias=bias)
    with np.errstate(all="ignore"):
        if inverse:
            if u[0] == 0:
                print("Warning: singular inverse transform; consider changing the bias", file=sys.stderr)
                u = u.copy()
                u[0] = np.inf
            u = 1.0 / np.conj(u)
        elif np.isinf(u[0]):
            print("Warning: singular transform; consider changing the bias", file=sys.stderr)
            u = u.copy()
            u[0] = 0.0
        table = u * np.exp(-2j * np.pi * np.arange(n // 2 + 1) / n)
    table = np.where(np.isfinite(table), table, 0.0)
    spec = torch.empty(x.shape[:-1] + (n // 2 + 1,), dtype=cdt, device=x.device)
    last = [x.dim() - 1]
    _ll.r2c(work, spec, last, True, 1.0)
    _ll.scale_lines(spec, torch.from_numpy(table.astype(np.complex64 if cdt == torch.complex64 else np.complex128)).to(x.device))
    out = work if owned else torch.empty_like(x)
    _ll.c2r(spec, out, last, True, 1.0 / n)
    if bias != 0.0:
        post = 1.0 / np.exp(-bias * (j - jc) * dln) if inverse else np.exp(-bias * ((j - jc) * dln + offset))
        _ll.scale_lines(out, torch.from_numpy(post.astype(rdt)).to(x.device))
    out = out.to(tdt)
    return out.cpu().numpy() if host else out


def fht(a, dln, mu, offset=0.0, bias=0.0):
    """scipy.fft.fht: fast Hankel transform of order `mu` of the real array `a` (last axis), whose samples
    are spaced uniformly in ln r with step `dln`.  NumPy in -> NumPy out (staged); CUDA tensor in -> tensor out."""
    return _hankel(a, dln, mu, offset, bias, inverse=False)


def ifht(A, dln, mu, offset=0.0, bias=0.0):
    """scipy.fft.ifht: inverse of `fht`."""
    return _hankel(A, dln, mu, offset, bias, inverse=True)
