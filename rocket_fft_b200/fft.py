"""numpy.fft / scipy.fft style functions on top of the low-level API -- the layer directly above
the transform path (reference: rocket_fft/overloads.py).

Argument handling follows the reference's host helpers, re-implemented here:
  * shape/axes normalisation (`s`/`n`, `axes`, negative axes; O:421-489), zero-pad-or-crop
    (O:575-609), dtype promotion to float32/float64/complex64/complex128 (O:48-64);
  * `norm` -> `fct` (O:513-551), with the 2(N+delta) factors of DCT/DST (O:505-510, 1012-1013);
  * DCT/DST type inversion 2<->3 for the inverse transforms (O:704-712), `orthogonalize` (O:715-732);
  * real input to fft/fftn goes through c2c_sym (O:958-968), irfft*/hfft through c2r with the
    output length rule n = 2(m-1) unless given (O:644-657).
Inputs may be NumPy arrays (host path) or torch CUDA tensors (device path, no copies beyond
pad/crop); the result lives where the input lives.
"""
from __future__ import annotations

import math

import numpy as np

from . import lowlevel as _ll


# ---------------------------------------------------------------------------------------------
# tiny array-backend shim (numpy or torch)
# ---------------------------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _np_dtype(x):
    if _is_torch(x):
        return np.dtype(str(x.dtype).replace("torch.", ""))
    return np.asarray(x).dtype


def _torch_dtype(dt):
    import torch

    return getattr(torch, np.dtype(dt).name)


def _as_dtype(x, dt):
    if _is_torch(x):
        return x if _np_dtype(x) == dt else x.to(_torch_dtype(dt))
    x = np.asarray(x)
    return x if x.dtype == dt else x.astype(dt)


def _empty(like, shape, dt):
    if _is_torch(like):
        import torch

        return torch.empty(tuple(shape), dtype=_torch_dtype(dt), device=like.device)
    return np.empty(tuple(shape), dtype=dt)


def _zeros(like, shape, dt):
    out = _empty(like, shape, dt)
    if _is_torch(out):
        out.zero_()
    else:
        out[...] = 0
    return out


def _real_of(dt):
    dt = np.dtype(dt)
    if dt in (np.dtype(np.complex64), np.dtype(np.float32), np.dtype(np.float16)):
        return np.dtype(np.float32)
    return np.dtype(np.float64)


def _cplx_of(dt):
    return np.dtype(np.complex64) if _real_of(dt) == np.dtype(np.float32) else np.dtype(np.complex128)


def _is_complex(dt):
    return np.dtype(dt).kind == "c"


# ---------------------------------------------------------------------------------------------
# NumPy / SciPy flavour (reference: rocket_fft/overloads.py:325-357, 380-405, rocket_fft/__init__.py:12-15)
# ---------------------------------------------------------------------------------------------
# SciPy rejects duplicate axes in the multi-axis transforms, NumPy accepts them (the axis is then transformed twice).
# Like the reference the SciPy behaviour is the default when SciPy is installed.
import importlib.util as _ilu

_unique_axes = _ilu.find_spec("scipy") is not None
_num_workers = 1


def numpy_like():
    """Duplicate axes are allowed (numpy.fft semantics)."""
    global _unique_axes
    _unique_axes = False


def scipy_like():
    """Duplicate axes raise ValueError("All axes must be unique.") (scipy.fft semantics)."""
    global _unique_axes
    _unique_axes = True


def get_workers():
    """The reference's worker count (O:345-346).  Accepted for interface compatibility: the parallelism here is the GPU grid."""
    return _num_workers


def set_workers(workers):
    """Validated like the reference (O:349-356) and otherwise ignored (`nthreads` has no meaning on the GPU)."""
    import os

    global _num_workers
    workers = int(workers)
    if workers < 1:
        raise ValueError("Number of workers cannot be smaller than one.")
    if workers > (os.cpu_count() or 1):
        raise ValueError(f"Number of workers exceeds CPU count of {os.cpu_count() or 1}.")
    _num_workers = workers


def _scipy_extras(overwrite_x, workers, plan=None):
    """The scipy.fft keyword arguments that have no meaning here, validated like the reference (O:555-572; `plan` as SciPy):
    `overwrite_x` is a permission, never a requirement (the result is always a new array); `workers` is checked and ignored."""
    import os

    if plan is not None:
        raise NotImplementedError("Passing a precomputed plan is not yet supported by scipy.fft functions")
    if workers is not None:
        workers = int(workers)
        if workers == 0:
            raise ValueError("Workers must not be zero.")
        if workers < -(os.cpu_count() or 1):
            raise ValueError("Workers value out of range.")


# ---------------------------------------------------------------------------------------------
# argument normalisation
# ---------------------------------------------------------------------------------------------
def _arr(x):
    """Array-likes that numpy / scipy accept (lists, tuples, scalars) become NumPy arrays; arrays and tensors pass."""
    return x if hasattr(x, "shape") and hasattr(x, "dtype") else np.asarray(x)


def _shape_axes(x, s, axes, default_all):
    nd = len(x.shape)
    if axes is None:
        if s is None:
            axes = list(range(nd)) if default_all else [nd - 1]
        else:
            if len(s if hasattr(s, "__len__") else [s]) > nd:
                raise ValueError("Shape requires more axes than are present.")
            axes = list(range(nd - len(s), nd))
    else:
        axes = [int(a) for a in (axes if hasattr(axes, "__len__") else [axes])]
    axes = [a + nd if a < 0 else a for a in axes]
    for a in axes:
        if not 0 <= a < nd:
            raise ValueError("axes exceeds dimensionality of input")
    if _unique_axes and len(set(axes)) != len(axes):
        raise ValueError("All axes must be unique.")
    if s is None:
        s = [x.shape[a] for a in axes]
    else:
        s = [int(v) for v in (s if hasattr(s, "__len__") else [s])]
        if len(s) != len(axes):
            raise ValueError("When given, axes and shape arguments have to be of the same length.")
        s = [x.shape[a] if v == -1 else v for v, a in zip(s, axes)]
    for v in s:
        if v < 1:
            raise ValueError(f"invalid number of data points ({v}) specified")
    return s, axes


def _target_shape(x, s, axes):
    shape = list(x.shape)
    for v, a in zip(s, axes):
        shape[a] = v
    return shape


def _pad_or_crop(x, s, axes, dt):
    """Host path / transforms without a fused-padding entry point: cropping is a view, zero-padding a copy
    (what the reference does for every call, O:575-609)."""
    x = _as_dtype(x, dt)
    shape = _target_shape(x, s, axes)
    if shape == list(x.shape):
        return x
    sl = tuple(slice(0, min(a, b)) for a, b in zip(shape, x.shape))
    if all(a <= b for a, b in zip(shape, x.shape)):
        return x[sl]
    out = _zeros(x, shape, dt)
    out[sl] = x[sl]
    return out


def _needs_pad(x, shape):
    return any(a > b for a, b in zip(shape, x.shape))


def _fct(shape, axes, norm, forward, delta=None):
    n = 1.0
    for a in axes:
        n *= shape[a] if delta is None else 2.0 * (shape[a] + delta)
    if norm is None or norm == "backward":
        return 1.0 if forward else 1.0 / n
    if norm == "ortho":
        return 1.0 / math.sqrt(n)
    if norm == "forward":
        return 1.0 / n if forward else 1.0
    raise ValueError("Invalid norm value; should be 'backward', 'ortho' or 'forward'.")


# ---------------------------------------------------------------------------------------------
# complex / real transforms
# ---------------------------------------------------------------------------------------------
def _c2cn(x, s, axes, norm, forward, default_all):
    x = _arr(x)
    s, axes = _shape_axes(x, s, axes, default_all)
    dt = _np_dtype(x)
    cdt = _cplx_of(dt)
    if _is_complex(dt):
        shape = _target_shape(x, s, axes)
        if _is_torch(x) and _needs_pad(x, shape):
            # device arrays: zero-padding is part of the first load of every line (rfb200_c2c_pad)
            out = _empty(x, shape, cdt)
            _ll.c2c_pad(_as_dtype(x, cdt), out, axes, forward, _fct(shape, axes, norm, forward))
            return out
        x = _pad_or_crop(x, s, axes, cdt)
        out = _empty(x, x.shape, cdt)
        _ll.c2c(x, out, axes, forward, _fct(x.shape, axes, norm, forward))
    else:
        x = _pad_or_crop(x, s, axes, _real_of(dt))
        out = _empty(x, x.shape, cdt)
        _ll.c2c_sym(x, out, axes, forward, _fct(x.shape, axes, norm, forward))
    return out


def _r2cn(x, s, axes, norm, forward, default_all):
    x = _arr(x)
    dt = _np_dtype(x)
    if _is_complex(dt):
        raise TypeError(f"unsupported dtype {dt}")
    s, axes = _shape_axes(x, s, axes, default_all)
    rdt = _real_of(dt)
    full = _target_shape(x, s, axes)
    if _is_torch(x) and _needs_pad(x, full):
        shape = list(full)
        shape[axes[-1]] = full[axes[-1]] // 2 + 1
        out = _empty(x, shape, _cplx_of(rdt))
        _ll.r2c_pad(_as_dtype(x, rdt), out, full, axes, forward, _fct(full, axes, norm, forward))
        return out
    x = _pad_or_crop(x, s, axes, rdt)
    shape = list(x.shape)
    shape[axes[-1]] = shape[axes[-1]] // 2 + 1
    out = _empty(x, shape, _cplx_of(rdt))
    _ll.r2c(x, out, axes, forward, _fct(x.shape, axes, norm, forward))
    return out


def _c2rn(x, s, axes, norm, forward, default_all):
    x = _arr(x)
    s_given = s is not None
    s, axes = _shape_axes(x, s, axes, default_all)
    cdt = _cplx_of(_np_dtype(x))
    last = axes[-1]
    n_last = s[-1] if s_given else 2 * (x.shape[last] - 1)
    if n_last < 1:
        raise ValueError(f"Invalid number of data points ({n_last}) specified")
    s_in = list(s)
    s_in[-1] = n_last // 2 + 1
    if _is_torch(x) and _needs_pad(x, _target_shape(x, s_in, axes)):
        shape = _target_shape(x, s, axes)
        shape[last] = n_last
        out = _empty(x, shape, _real_of(cdt))
        _ll.c2r_pad(_as_dtype(x, cdt), out, axes, forward, _fct(shape, axes, norm, forward))
        return out
    xin = _pad_or_crop(x, s_in, axes, cdt)
    shape = list(xin.shape)
    shape[last] = n_last
    out = _empty(xin, shape, _real_of(cdt))
    _ll.c2r(xin, out, axes, forward, _fct(shape, axes, norm, forward))
    return out


def fft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2cn(x, None if n is None else [n], [axis], norm, True, False)


def ifft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2cn(x, None if n is None else [n], [axis], norm, False, False)


def fft2(x, s=None, axes=(-2, -1), norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2cn(x, s, axes, norm, True, True)


def ifft2(x, s=None, axes=(-2, -1), norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2cn(x, s, axes, norm, False, True)


def fftn(x, s=None, axes=None, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2cn(x, s, axes, norm, True, True)


def ifftn(x, s=None, axes=None, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2cn(x, s, axes, norm, False, True)


def rfft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _r2cn(x, None if n is None else [n], [axis], norm, True, False)


def irfft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2rn(x, None if n is None else [n], [axis], norm, False, False)


def rfft2(x, s=None, axes=(-2, -1), norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _r2cn(x, s, axes, norm, True, True)


def irfft2(x, s=None, axes=(-2, -1), norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2rn(x, s, axes, norm, False, True)


def rfftn(x, s=None, axes=None, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _r2cn(x, s, axes, norm, True, True)


def irfftn(x, s=None, axes=None, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _c2rn(x, s, axes, norm, False, True)


def hfft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    """FFT of a Hermitian-symmetric signal (real spectrum): c2r with the forward sign."""
    _scipy_extras(overwrite_x, workers, plan)
    return _c2rn(x, None if n is None else [n], [axis], norm, True, False)


def ihfft(x, n=None, axis=-1, norm=None, overwrite_x=False, workers=None, *, plan=None):
    _scipy_extras(overwrite_x, workers, plan)
    return _r2cn(x, None if n is None else [n], [axis], norm, False, False)


def hfft2(x, s=None, axes=(-2, -1), norm=None, overwrite_x=False, workers=None, *, plan=None):
    """scipy.fft.hfft2 (O:1516-1528): c2r over the given axes with the forward sign."""
    _scipy_extras(overwrite_x, workers, plan)
    return _c2rn(x, s, axes, norm, True, True)


def ihfft2(x, s=None, axes=(-2, -1), norm=None, overwrite_x=False, workers=None, *, plan=None):
    """scipy.fft.ihfft2 (O:1530-1542): r2c over the given axes with the backward sign."""
    _scipy_extras(overwrite_x, workers, plan)
    return _r2cn(x, s, axes, norm, False, True)


def hfftn(x, s=None, axes=None, norm=None, overwrite_x=False, workers=None, *, plan=None):
    """scipy.fft.hfftn (O:1544-1556)."""
    _scipy_extras(overwrite_x, workers, plan)
    return _c2rn(x, s, axes, norm, True, True)


def ihfftn(x, s=None, axes=None, norm=None, overwrite_x=False, workers=None, *, plan=None):
    """scipy.fft.ihfftn (O:1558-1570)."""
    _scipy_extras(overwrite_x, workers, plan)
    return _r2cn(x, s, axes, norm, False, True)


# ---------------------------------------------------------------------------------------------
# DCT / DST
# ---------------------------------------------------------------------------------------------
def _r2rn(x, type, s, axes, norm, orthogonalize, forward, cosine, default_all):
    x = _arr(x)
    type = int(type)
    if type not in (1, 2, 3, 4):
        raise ValueError("Invalid type; must be one of (1, 2, 3, 4).")
    s, axes = _shape_axes(x, s, axes, default_all)
    dt = _np_dtype(x)
    if not forward:
        type = {1: 1, 2: 3, 3: 2, 4: 4}[type]
    delta = (-1.0 if cosine else 1.0) if type == 1 else 0.0
    ortho = (norm == "ortho") if orthogonalize is None else bool(orthogonalize)
    if cosine:
        fn = _ll.dct
    elif ortho and type in (2, 3):
        # this layer follows scipy.fft: DST-II/III under ortho scale element N-1 (the reference's low-level dst scales
        # element 0 -- a known quirk it warns about, README.md:61-65; the numba_* drop-in symbols keep it)
        def fn(a, o, ax, t, f, orth):
            return _ll.dst(a, o, ax, t, f, orth, dst_ortho="scipy")
    else:
        fn = _ll.dst
    if _is_complex(dt):
        cdt = _cplx_of(dt)
        x = _pad_or_crop(x, s, axes, cdt)
        out = _empty(x, x.shape, cdt)
        fct = _fct(out.shape, axes, norm, forward, delta)
        for part in ("real", "imag"):
            xi, oi = getattr(x, part), getattr(out, part)
            if not _is_torch(x):
                fn(xi, oi, axes, type, fct, ortho)
            else:
                import torch

                xr = torch.view_as_real(x)[..., 0 if part == "real" else 1]
                orr = torch.view_as_real(out)[..., 0 if part == "real" else 1]
                fn(xr, orr, axes, type, fct, ortho)
        return out
    rdt = _real_of(dt)
    x = _pad_or_crop(x, s, axes, rdt)
    out = _empty(x, x.shape, rdt)
    fn(x, out, axes, type, _fct(out.shape, axes, norm, forward, delta), ortho)
    return out


def dct(x, type=2, n=None, axis=-1, norm=None, orthogonalize=None, *, overwrite_x=False, workers=None):
    _scipy_extras(overwrite_x, workers)
    return _r2rn(x, type, None if n is None else [n], [axis], norm, orthogonalize, True, True, False)


def idct(x, type=2, n=None, axis=-1, norm=None, orthogonalize=None, *, overwrite_x=False, workers=None):
    _scipy_extras(overwrite_x, workers)
    return _r2rn(x, type, None if n is None else [n], [axis], norm, orthogonalize, False, True, False)


def dst(x, type=2, n=None, axis=-1, norm=None, orthogonalize=None, *, overwrite_x=False, workers=None):
    _scipy_extras(overwrite_x, workers)
    return _r2rn(x, type, None if n is None else [n], [axis], norm, orthogonalize, True, False, False)


def idst(x, type=2, n=None, axis=-1, norm=None, orthogonalize=None, *, overwrite_x=False, workers=None):
    _scipy_extras(overwrite_x, workers)
    return _r2rn(x, type, None if n is None else [n], [axis], norm, orthogonalize, False, False, False)


def dctn(x, type=2, s=None, axes=None, norm=None, orthogonalize=None, *, overwrite_x=False, workers=None):
    _scipy_extras(overwrite_x, workers)
    return _r2rn(x, type, s, axes, norm, orthogonalize, True, True, True)


def idctn(x, type=2, s=None, axes=None, norm=None, orthogonalize=None, *, overwrite_x=False, workers=None):
    _scipy_extras(overwrite_x, workers)
    return _r2rn(x, type, s, axes, norm, orthogonalize, False, True, True)


def dstn(x, type=2, s=None, axes=None, norm=None, orthogonalize=None, *, overwrite_x=False, workers=None):
    _scipy_extras(overwrite_x, workers)
    return _r2rn(x, type, s, axes, norm, orthogonalize, True, False, True)


def idstn(x, type=2, s=None, axes=None, norm=None, orthogonalize=None, *, overwrite_x=False, workers=None):
    _scipy_extras(overwrite_x, workers)
    return _r2rn(x, type, s, axes, norm, orthogonalize, False, False, True)


def next_fast_len(target, real=False):
    """scipy.fft.next_fast_len (O:1862-1870 -> numba_good_size)."""
    if target < 0:
        raise ValueError("Target cannot be negative.")
    return _ll.good_size(int(target), bool(real))


def prev_fast_len(target, real=False):
    """scipy.fft.prev_fast_len: the largest 11-smooth (complex) or 5-smooth (real) length <= target -- the companion of
    next_fast_len for trimming data.  (The reference leaves it as a TODO, O:1872-1877; host-only integer code, same
    factor sets as good_size, H:589-650.)  target == 0 returns 0 like SciPy; negative targets raise."""
    target = int(target)
    if target < 0:
        raise ValueError("Target length must be positive")
    if target <= (6 if real else 12):
        return target
    primes = (2, 3, 5) if real else (2, 3, 5, 7, 11)
    best = 1

    def search(i, prod):
        nonlocal best
        if i == len(primes):
            best = max(best, prod)
            return
        p = primes[i]
        while prod <= target:
            search(i + 1, prod)
            prod *= p

    search(0, 1)
    return best


# ---------------------------------------------------------------------------------------------
# helpers either side of a transform (reference: rocket_fft/overloads.py:752-855, 1221-1310)
# ---------------------------------------------------------------------------------------------
def _roll_axes(x, axes, sign):
    nd = len(x.shape)
    if axes is None:
        axes = list(range(nd))
    elif not hasattr(axes, "__len__"):
        axes = [int(axes)]
    shift = [0] * nd
    for a in axes:
        a = int(a)
        if not -nd <= a < nd:
            raise ValueError("axes exceeds dimensionality of input")
        shift[a % nd] += sign * (x.shape[a % nd] // 2)
    return roll(x, shift, axis=list(range(nd)))


def roll(x, shift, axis=None):
    """numpy.roll: a single pass of the index-rotation kernel (rfb200_roll).  NumPy inputs are staged through
    device memory (H2D, kernel, D2H); torch CUDA tensors stay on the device."""
    import torch

    host = not _is_torch(x)
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda() if host else x
    nd = xd.dim()
    if axis is None:
        flat = xd.reshape(-1)
        out = torch.empty_like(flat)
        total = int(np.sum(shift)) if hasattr(shift, "__len__") else int(shift)
        _roll_any(flat, out, [total])
        out = out.reshape(xd.shape)
    else:
        axes = [int(a) for a in (axis if hasattr(axis, "__len__") else [axis])]
        shifts = [int(v) for v in (shift if hasattr(shift, "__len__") else [shift] * len(axes))]
        if len(shifts) != len(axes):
            raise ValueError("shift and axis must have the same length")
        per = [0] * nd
        for a, v in zip(axes, shifts):
            if not -nd <= a < nd:
                raise ValueError("axis out of range")
            per[a % nd] += v
        out = torch.empty(xd.shape, dtype=xd.dtype, device=xd.device)
        _roll_any(xd, out, per)
    return out.cpu().numpy() if host else out


def _roll_any(x, out, per):
    import torch

    item = x.element_size()
    if item in (4, 8, 16):
        _ll.roll(x, out, per)
        return
    if x.numel() == 0:
        return
    # 1- and 2-byte items: move them as 4-byte words when the last dim allows it, else widen
    wide = x.to(torch.float32 if x.is_floating_point() else torch.int32)
    tmp = torch.empty_like(wide)
    _ll.roll(wide, tmp, per)
    out.copy_(tmp.to(x.dtype))


def fftshift(x, axes=None):
    """Zero-frequency bin to the centre: out[(i + n//2) mod n] = x[i] along `axes` (all by default)."""
    return _roll_axes(x, axes, +1)


def ifftshift(x, axes=None):
    """Inverse of fftshift: out[(i - n//2) mod n] = x[i]."""
    return _roll_axes(x, axes, -1)


def fftfreq(n, d=1.0, device=None):
    """Sample frequencies [0, 1, ..., (n-1)//2, -(n//2), ..., -1] / (d n) (float64); on `device` if given."""
    n = int(n)
    if n < 1:
        raise ValueError(f"n should be an integer greater than 0, got {n}")
    k = np.arange(n, dtype=np.int64)
    k[(n - 1) // 2 + 1:] -= n
    out = k * (1.0 / (n * d))  # same rounding as numpy: multiply by the reciprocal
    if device is not None:
        import torch

        return torch.from_numpy(out).to(device)
    return out


def rfftfreq(n, d=1.0, device=None):
    """Sample frequencies [0, 1, ..., n//2] / (d n) (float64); on `device` if given."""
    n = int(n)
    if n < 1:
        raise ValueError(f"n should be an integer greater than 0, got {n}")
    out = np.arange(n // 2 + 1, dtype=np.int64) * (1.0 / (n * d))
    if device is not None:
        import torch

        return torch.from_numpy(out).to(device)
    return out
