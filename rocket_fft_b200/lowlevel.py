"""Low-level transform API of rocket_fft_b200 -- same names, argument order and
meaning as the reference's low-level interface (rocket_fft/__init__.py:18-34,
rocket_fft/__init__.pyi:6-105; README.md:74-85):

    c2c(ain, aout, axes, forward, fct, nthreads)        r2c / c2r / c2c_sym likewise
    dct(ain, aout, axes, type, fct, ortho, nthreads)    dst likewise
    r2r_separable_hartley(ain, aout, axes, fct, nthreads)   r2r_genuine_hartley likewise
    r2r_fftpack(ain, aout, axes, real2hermitian, forward, fct, nthreads)
    good_size(n, real)

Arrays may be
  * NumPy arrays (host): the call goes through the library's ``numba_*`` symbols --
    the very entry points Numba-compiled code binds -- which stage H2D, run the
    sm_100a kernels and copy the result back before returning;
  * anything exporting ``__cuda_array_interface__`` (torch CUDA tensors, CuPy, Numba
    device arrays): the ``rfb200_*`` device entry points run on the caller's current
    CUDA stream with no copies.
``nthreads`` is accepted for signature compatibility and ignored: the parallelism is
the GPU's.  There is no CPU fallback; importing this module without the compiled
library raises ImportError.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librocketfft_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C rocket_fft_b200/csrc` (there is no CPU fallback)"
    )

lib = _abi.LowLevelLib(LIB_PATH)
_c = lib.cdll

_I64P = C.POINTER(C.c_int64)
_U64P = C.POINTER(C.c_uint64)
_DEV_COMMON = [C.c_int, C.c_size_t, _I64P, _I64P, _I64P, C.c_size_t, _U64P]
_DEV_TAIL = [C.c_void_p, C.c_void_p, C.c_void_p]
for _name, _mid in {
    "rfb200_c2c": [C.c_int, C.c_double],
    "rfb200_r2c": [C.c_int, C.c_double],
    "rfb200_c2r": [C.c_int, C.c_double],
    "rfb200_c2c_sym": [C.c_int, C.c_double],
    "rfb200_dct": [C.c_int, C.c_double, C.c_int],
    "rfb200_dst": [C.c_int, C.c_double, C.c_int],
    "rfb200_r2r_fftpack": [C.c_int, C.c_int, C.c_double],
    "rfb200_r2r_separable_hartley": [C.c_double],
    "rfb200_r2r_genuine_hartley": [C.c_double],
}.items():
    _f = getattr(_c, _name)
    _f.restype = C.c_int
    _f.argtypes = _DEV_COMMON + _mid + _DEV_TAIL
_c.rfb200_last_error.restype = C.c_char_p
_c.rfb200_version.restype = C.c_char_p
_c.rfb200_launch_count.restype = C.c_uint64
_c.rfb200_set_stream.argtypes = [C.c_void_p]
_c.rfb200_set_dst_ortho_quirk.argtypes = [C.c_int]
_c.rfb200_failure_count.restype = C.c_uint64
_c.rfb200_plan_cache_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
_c.rfb200_host_dst.restype = None
_c.rfb200_host_dst.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_int]


class TransformError(RuntimeError):
    pass


def last_error() -> str:
    return (_c.rfb200_last_error() or b"").decode()


def version() -> str:
    return _c.rfb200_version().decode()


def launch_count() -> int:
    return int(_c.rfb200_launch_count())


def launch_count_reset() -> None:
    _c.rfb200_launch_count_reset()


_c.rfb200_launch_trace.argtypes = [C.c_int]
_c.rfb200_launch_trace_get.restype = C.c_char_p


def launch_trace(enable: bool) -> None:
    """Start (and clear) / stop the record of kernel names launched by the library."""
    _c.rfb200_launch_trace(1 if enable else 0)


def launch_trace_get():
    """Names of the kernels launched since launch_trace(True), in launch order."""
    s = (_c.rfb200_launch_trace_get() or b"").decode()
    return s.split(";") if s else []


def plan_cache_clear() -> None:
    _c.rfb200_plan_cache_clear()


def set_dst_ortho_quirk(enabled: bool) -> None:
    """True (default): DST-II/III with ortho=True scale element 0 like the reference
    (README.md:61-65); False: element N-1 like SciPy."""
    _c.rfb200_set_dst_ortho_quirk(1 if enabled else 0)


def set_stream(stream) -> None:
    """CUDA stream (integer handle) used by the host-array path of this thread;
    None restores the library's own stream."""
    if stream is None:
        _c.rfb200_use_library_stream()
    else:
        _c.rfb200_set_stream(C.c_void_p(int(stream)))


def failure_count() -> int:
    """Number of numba_* calls that failed since the library was loaded (those entry points return void: a failed call
    prints its reason, fills a host output with NaN and bumps this counter)."""
    return int(_c.rfb200_failure_count())


_c.rfb200_host_pin.restype = C.c_int
_c.rfb200_host_pin.argtypes = [C.c_void_p, C.c_uint64]
_c.rfb200_host_unpin.restype = C.c_int
_c.rfb200_host_unpin.argtypes = [C.c_void_p]


class pinned:
    """Context manager: page-locks the buffers of the given C-contiguous NumPy arrays for its duration (rfb200_host_pin), so
    that host-array calls on them copy at the rate of pinned memory instead of going through the staging ring:

        with rfb.pinned(x, out):
            for _ in range(steps):
                rfb.r2c(x, out, [1, 2], True, 1.0)

    Pinning costs about 0.1 s per GiB (measured on the B200 box), so it pays for buffers that are used more than once.
    The arrays must stay alive (and not be resized) inside the block."""

    def __init__(self, *arrays):
        self.arrays = arrays
        self.done = []

    def __enter__(self):
        for a in self.arrays:
            if not isinstance(a, np.ndarray) or not a.flags.c_contiguous:
                raise TypeError("pinned() takes C-contiguous NumPy arrays")
            if a.nbytes == 0:
                continue
            if _c.rfb200_host_pin(C.c_void_p(a.ctypes.data), a.nbytes) != 0:
                err = last_error()
                self.__exit__(None, None, None)
                raise TransformError(err)
            self.done.append(a.ctypes.data)
        return self

    def __exit__(self, *exc):
        for p in self.done:
            _c.rfb200_host_unpin(C.c_void_p(p))
        self.done = []
        return False


def plan_cache_stats():
    """(entries, bytes) of the device table cache (all devices)."""
    e, b = C.c_uint64(0), C.c_uint64(0)
    _c.rfb200_plan_cache_stats(C.byref(e), C.byref(b))
    return int(e.value), int(b.value)


def _current_stream(a=None) -> int:
    """The caller's current stream ON THE DEVICE THAT HOLDS `a` (torch tensors; other exporters: the current device's)."""
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available():
        dev = getattr(a, "device", None)
        if dev is not None and getattr(dev, "type", None) == "cuda":
            return int(torch.cuda.current_stream(dev).cuda_stream)
        return int(torch.cuda.current_stream().cuda_stream)
    return 0


def _same_device(*arrays):
    devs = {str(a.device) for a in arrays if getattr(getattr(a, "device", None), "type", None) == "cuda"}
    if len(devs) > 1:
        raise ValueError(f"arrays live on different devices: {sorted(devs)}")


def _is_host(a) -> bool:
    return isinstance(a, np.ndarray)


_REAL = (np.dtype(np.float32), np.dtype(np.float64))
_CPLX = (np.dtype(np.complex64), np.dtype(np.complex128))


def _check(op, ain, aout):
    din, dout = _abi.dtype_of(ain), _abi.dtype_of(aout)
    want_in = _CPLX if op in ("c2c", "c2r") else _REAL
    want_out = _CPLX if op in ("c2c", "r2c", "c2c_sym") else _REAL
    if din not in want_in or dout not in want_out:
        raise TypeError(f"{op}: unsupported dtypes {din} -> {dout}")
    if (din in (np.dtype(np.float32), np.dtype(np.complex64))) != (dout in (np.dtype(np.float32), np.dtype(np.complex64))):
        raise TypeError(f"{op}: input and output precision differ ({din} -> {dout})")
    return 0 if din in (np.dtype(np.float32), np.dtype(np.complex64)) else 1


def _device_call(op, ain, aout, axes, mid):
    prec = _check(op, ain, aout)
    pin, shp_in, st_in, _, dev_in, k1 = _abi.describe(ain)
    pout, shp_out, st_out, _, dev_out, k2 = _abi.describe(aout)
    if len(shp_in) != len(shp_out):
        raise ValueError("Input and output array must have the same number of dimensions")
    shape = shp_out if op == "c2r" else shp_in
    nd = len(shape)
    ax = [int(a) for a in np.asarray(axes).ravel()]
    for a in ax:
        if not 0 <= a < nd:
            raise ValueError("axis out of range")
    A = (C.c_int64 * max(nd, 1))
    shape_c, sin_c, sout_c = A(*shape), A(*st_in), A(*st_out)
    axes_c = (C.c_uint64 * max(len(ax), 1))(*ax)
    _same_device(ain, aout)
    rc = getattr(_c, "rfb200_" + op)(
        prec, nd, shape_c, sin_c, sout_c, len(ax), axes_c, *mid, C.c_void_p(pin), C.c_void_p(pout),
        C.c_void_p(_current_stream(ain)),
    )
    if rc != 0:
        raise TransformError(last_error())
    return aout


def _host_call(op, ain, aout, axes, args):
    _check(op, ain, aout)
    getattr(lib, op)(ain, aout, axes, *args)
    err = last_error()
    if err:
        raise TransformError(err)
    return aout


def _call(op, ain, aout, axes, host_args, dev_mid):
    hi, ho = _is_host(ain), _is_host(aout)
    if hi != ho:
        raise TypeError("input and output must both be host arrays or both be device arrays")
    if hi:
        return _host_call(op, ain, aout, axes, host_args)
    return _device_call(op, ain, aout, axes, dev_mid)


def c2c(ain, aout, axes, forward, fct, nthreads=1):
    return _call("c2c", ain, aout, axes, (forward, fct, nthreads), (int(bool(forward)), float(fct)))


def r2c(ain, aout, axes, forward, fct, nthreads=1):
    return _call("r2c", ain, aout, axes, (forward, fct, nthreads), (int(bool(forward)), float(fct)))


def c2r(ain, aout, axes, forward, fct, nthreads=1):
    return _call("c2r", ain, aout, axes, (forward, fct, nthreads), (int(bool(forward)), float(fct)))


def c2c_sym(ain, aout, axes, forward, fct, nthreads=1):
    return _call("c2c_sym", ain, aout, axes, (forward, fct, nthreads), (int(bool(forward)), float(fct)))


def dct(ain, aout, axes, type, fct, ortho, nthreads=1):
    if int(type) not in (1, 2, 3, 4):
        raise ValueError("invalid DCT type")
    return _call("dct", ain, aout, axes, (type, fct, ortho, nthreads), (int(type), float(fct), int(bool(ortho))))


def dst(ain, aout, axes, type, fct, ortho, nthreads=1, *, dst_ortho=None):
    """`dst_ortho` chooses the DST-II/III scaling under ortho=True for THIS call: None = the process-wide setting
    (set_dst_ortho_quirk; default: the reference's, which scales element 0, README.md:61-65), "scipy" = element N-1 as
    SciPy does, "reference" = the reference's."""
    if int(type) not in (1, 2, 3, 4):
        raise ValueError("invalid DST type")
    if dst_ortho not in (None, "scipy", "reference"):
        raise ValueError("dst_ortho must be None, 'scipy' or 'reference'")
    if dst_ortho is None or not ortho:
        return _call("dst", ain, aout, axes, (type, fct, ortho, nthreads), (int(type), float(fct), int(bool(ortho))))
    if _is_host(ain) != _is_host(aout):
        raise TypeError("input and output must both be host arrays or both be device arrays")
    if not _is_host(ain):
        return _device_call("dst", ain, aout, axes, (int(type), float(fct), 2 if dst_ortho == "scipy" else 3))
    _check("dst", ain, aout)
    nd = len(_abi.describe(ain)[1])
    if len(_abi.describe(aout)[1]) != nd:
        raise ValueError("Input and output array must have the same number of dimensions")
    rin, k1 = _abi.make_record(ain)
    rout, k2 = (rin, k1) if aout is ain else _abi.make_record(aout)
    rax, k3 = _abi.axes_record(axes)
    _c.rfb200_host_dst(nd, C.addressof(rin), C.addressof(rout), C.addressof(rax), int(type), float(fct), 1,
                       0 if dst_ortho == "scipy" else 1)
    err = last_error()
    if err:
        raise TransformError(err)
    return aout


def r2r_separable_hartley(ain, aout, axes, fct, nthreads=1):
    return _call("r2r_separable_hartley", ain, aout, axes, (fct, nthreads), (float(fct),))


def r2r_genuine_hartley(ain, aout, axes, fct, nthreads=1):
    return _call("r2r_genuine_hartley", ain, aout, axes, (fct, nthreads), (float(fct),))


def r2r_fftpack(ain, aout, axes, real2hermitian, forward, fct, nthreads=1):
    return _call(
        "r2r_fftpack", ain, aout, axes, (real2hermitian, forward, fct, nthreads),
        (int(bool(real2hermitian)), int(bool(forward)), float(fct)),
    )


_c.rfb200_c2c_scatter.restype = C.c_int
_c.rfb200_c2c_scatter.argtypes = [C.c_int, C.c_size_t, _I64P, _I64P, _I64P, C.c_size_t, C.c_int, C.c_double,
                                  C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]


def c2c_scatter(ain, parts, axis, forward, fct):
    """c2c along `axis` of the device array `ain`; the output index along that axis is cut into
    len(parts) equal blocks and block h is written into parts[h] (device arrays, possibly
    peer-mapped, all with the same strides and the axis extent n/len(parts))."""
    pin, shp, st_in, _, dev_in, _k = _abi.describe(ain)
    if not dev_in:
        raise TypeError("c2c_scatter works on device arrays")
    prec = 0 if _abi.dtype_of(ain) == np.dtype(np.complex64) else 1
    nd = len(shp)
    desc = [_abi.describe(p) for p in parts]
    st_out = desc[0][2]
    for d in desc:
        if d[2] != st_out or not d[4]:
            raise ValueError("all parts must be device arrays with identical strides")
    A = (C.c_int64 * nd)
    ptrs = (C.c_void_p * len(parts))(*[d[0] for d in desc])
    rc = _c.rfb200_c2c_scatter(prec, nd, A(*shp), A(*st_in), A(*st_out), int(axis) % nd, int(bool(forward)), float(fct),
                               C.c_void_p(pin), len(parts), ptrs, C.c_void_p(_current_stream(ain)))
    if rc != 0:
        raise TransformError(last_error())
    return parts


_PAD_ARGS = [C.c_int, C.c_size_t, _I64P, _I64P, _I64P, _I64P, C.c_size_t, _U64P, C.c_int, C.c_double,
             C.c_void_p, C.c_void_p, C.c_void_p]
for _name in ("rfb200_c2c_pad", "rfb200_r2c_pad", "rfb200_c2r_pad"):
    getattr(_c, _name).restype = C.c_int
    getattr(_c, _name).argtypes = _PAD_ARGS
_c.rfb200_roll.restype = C.c_int
_c.rfb200_roll.argtypes = [C.c_int, C.c_size_t, _I64P, _I64P, _I64P, _I64P, C.c_void_p, C.c_void_p, C.c_void_p]


def _pad_call(op, ain, aout, shape, axes, forward, fct):
    """Device arrays only: `ain` as it is, `shape` what is transformed (see include/rocketfft_b200.h (3))."""
    base = op[:-4]
    prec = _check(base, ain, aout)
    pin, shp_in, st_in, _, dev_in, _k1 = _abi.describe(ain)
    pout, shp_out, st_out, _, dev_out, _k2 = _abi.describe(aout)
    if not (dev_in and dev_out):
        raise TypeError(f"{op} works on device arrays")
    nd = len(shp_in)
    shape = tuple(int(v) for v in shape)
    if len(shp_out) != nd or len(shape) != nd:
        raise ValueError("Input, output and transform shape must have the same number of dimensions")
    ax = [int(a) for a in np.asarray(axes).ravel()]
    for a in ax:
        if not 0 <= a < nd:
            raise ValueError("axis out of range")
    want_out = list(shape)
    if base == "r2c" and ax:
        want_out[ax[-1]] = shape[ax[-1]] // 2 + 1
    if tuple(want_out) != tuple(shp_out):
        raise ValueError(f"{op}: output shape {tuple(shp_out)} does not match the transform shape {tuple(want_out)}")
    A = (C.c_int64 * max(nd, 1))
    rc = getattr(_c, "rfb200_" + op)(
        prec, nd, A(*shp_in), A(*shape), A(*st_in), A(*st_out), len(ax), (C.c_uint64 * max(len(ax), 1))(*ax),
        int(bool(forward)), float(fct), C.c_void_p(pin), C.c_void_p(pout), C.c_void_p(_current_stream(ain)),
    )
    if rc != 0:
        raise TransformError(last_error())
    return aout


def c2c_pad(ain, aout, axes, forward, fct):
    """c2c of `ain` cropped / zero-extended to aout's shape along `axes`, without a padded copy."""
    return _pad_call("c2c_pad", ain, aout, _abi.describe(aout)[1], axes, forward, fct)


def r2c_pad(ain, aout, shape, axes, forward, fct):
    """r2c of the real array `ain` cropped / zero-extended to `shape` (the real transform shape)."""
    return _pad_call("r2c_pad", ain, aout, shape, axes, forward, fct)


def c2r_pad(ain, aout, axes, forward, fct):
    """c2r producing `aout` (its shape is the transform shape) from the bins present in `ain`."""
    return _pad_call("c2r_pad", ain, aout, _abi.describe(aout)[1], axes, forward, fct)


def roll(ain, aout, shift):
    """aout[(i + shift[d]) mod n_d] = ain[i] along every dim (device arrays; 4/8/16-byte items)."""
    pin, shp_in, st_in, item, dev_in, _k1 = _abi.describe(ain)
    pout, shp_out, st_out, item2, dev_out, _k2 = _abi.describe(aout)
    if not (dev_in and dev_out):
        raise TypeError("roll works on device arrays")
    if tuple(shp_in) != tuple(shp_out) or item != item2:
        raise ValueError("roll: input and output must have the same shape and item size")
    nd = len(shp_in)
    sh = [int(v) for v in shift]
    if len(sh) != nd:
        raise ValueError("roll: one shift per dimension")
    A = (C.c_int64 * max(nd, 1))
    rc = _c.rfb200_roll(int(item), nd, A(*shp_in), A(*st_in), A(*st_out), A(*sh), C.c_void_p(pin), C.c_void_p(pout),
                        C.c_void_p(_current_stream(ain)))
    if rc != 0:
        raise TransformError(last_error())
    return aout


_c.rfb200_scale_lines.restype = C.c_int
_c.rfb200_scale_lines.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]


def scale_lines(data, table):
    """data[..., j] *= table[j] in place; `data` a contiguous device array, `table` a contiguous 1-D device array of
    the same dtype (float32/64 or complex64/128) and of data's last extent."""
    pd, shp, st, item, dev, _k1 = _abi.describe(data)
    pt, tshp, tst, titem, tdev, _k2 = _abi.describe(table)
    dt = _abi.dtype_of(data)
    if not (dev and tdev) or dt != _abi.dtype_of(table) or dt not in _REAL + _CPLX:
        raise TypeError("scale_lines: device arrays of one real or complex floating dtype expected")
    n = shp[-1] if shp else 1
    acc = item
    for ext, sb in zip(reversed(shp), reversed(st)):
        if ext > 1 and sb != acc:
            raise ValueError("scale_lines: data must be C-contiguous")
        acc *= max(ext, 1)
    if tuple(tshp) != (n,) or (n > 1 and tst[0] != titem):
        raise ValueError("scale_lines: table must be contiguous with data's last extent")
    total = 1
    for ext in shp:
        total *= ext
    if total == 0:
        return data
    prec = 0 if dt in (np.dtype(np.float32), np.dtype(np.complex64)) else 1
    rc = _c.rfb200_scale_lines(prec, int(dt in _CPLX), total // n, n, C.c_void_p(pt), C.c_void_p(pd),
                               C.c_void_p(_current_stream(data)))
    if rc != 0:
        raise TransformError(last_error())
    return data


def good_size(n, real):
    return lib.good_size(n, real)


separable_hartley = r2r_separable_hartley
genuine_hartley = r2r_genuine_hartley
fftpack = r2r_fftpack
