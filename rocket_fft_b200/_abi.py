"""ctypes view of the C ABI that sits between the Python/Numba host layer and the
transform library.

The ten ``numba_*`` entry points take pointers to Numba's array record
(reference: rocket_fft/_pocketfft_numba.cpp:31-49 reads ``shape_and_strides``,
``data``, ``itemsize`` and ``nitems`` out of it; the record layout itself is
numba's ``_arraystruct.h``).  This module builds such records from NumPy arrays
(host memory) or from anything exporting ``__cuda_array_interface__`` (device
memory), and binds the entry points of *a* shared library that exports them.

``LowLevelLib`` is deliberately library-agnostic: the product binds it to
``librocketfft_b200.so``; the test-suite binds a second instance to the compiled
reference in ``oracle/_ref`` so both sides are driven through byte-identical calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

c_intp = C.c_ssize_t


def _record_type(ndim: int):
    class ArrayRecord(C.Structure):
        _fields_ = [
            ("meminfo", C.c_void_p),
            ("parent", C.c_void_p),
            ("nitems", c_intp),
            ("itemsize", c_intp),
            ("data", C.c_void_p),
            ("shape_and_strides", c_intp * (2 * max(ndim, 1))),
        ]

    return ArrayRecord


_RECORD_TYPES = {}


def record_type(ndim: int):
    t = _RECORD_TYPES.get(ndim)
    if t is None:
        t = _RECORD_TYPES[ndim] = _record_type(ndim)
    return t


def describe(a):
    """Return (data_ptr, shape, strides_bytes, itemsize, is_device, keepalive)."""
    if isinstance(a, np.ndarray):
        return a.ctypes.data, tuple(a.shape), tuple(a.strides), a.itemsize, False, a
    cai = getattr(a, "__cuda_array_interface__", None)
    if cai is None:
        raise TypeError(
            "expected a numpy.ndarray or an object exporting __cuda_array_interface__, "
            f"got {type(a).__name__}"
        )
    shape = tuple(int(s) for s in cai["shape"])
    itemsize = np.dtype(cai["typestr"]).itemsize
    strides = cai.get("strides")
    if strides is None:
        strides = []
        acc = itemsize
        for s in reversed(shape):
            strides.append(acc)
            acc *= max(int(s), 1)
        strides = tuple(reversed(strides))
    else:
        strides = tuple(int(s) for s in strides)
    return int(cai["data"][0] or 0), shape, strides, itemsize, True, a


def dtype_of(a) -> np.dtype:
    if isinstance(a, np.ndarray):
        return a.dtype
    return np.dtype(a.__cuda_array_interface__["typestr"])


def make_record(a):
    """Build the array record for ``a``; returns (record, keepalive)."""
    ptr, shape, strides, itemsize, _dev, keep = describe(a)
    ndim = len(shape)
    rec = record_type(ndim)()
    rec.meminfo = None
    rec.parent = None
    n = 1
    for s in shape:
        n *= s
    rec.nitems = n
    rec.itemsize = itemsize
    rec.data = ptr
    for i in range(ndim):
        rec.shape_and_strides[i] = shape[i]
        rec.shape_and_strides[ndim + i] = strides[i]
    return rec, keep


def axes_record(axes):
    ax = np.ascontiguousarray(np.asarray(axes).astype(np.uint64, copy=False).ravel())
    rec, keep = make_record(ax)
    return rec, keep


_VP = C.c_void_p
_U64 = C.c_uint64
_SIGS = {
    # name: (restype, argtypes)  -- reference: rocket_fft/pocketfft.py:33-128
    "numba_good_size": (_U64, (_U64, C.c_bool)),
    "numba_c2c": (None, (_U64, _VP, _VP, _VP, C.c_bool, C.c_double, _U64)),
    "numba_r2c": (None, (_U64, _VP, _VP, _VP, C.c_bool, C.c_double, _U64)),
    "numba_c2r": (None, (_U64, _VP, _VP, _VP, C.c_bool, C.c_double, _U64)),
    "numba_c2c_sym": (None, (_U64, _VP, _VP, _VP, C.c_bool, C.c_double, _U64)),
    "numba_dct": (None, (_U64, _VP, _VP, _VP, _U64, C.c_double, C.c_bool, _U64)),
    "numba_dst": (None, (_U64, _VP, _VP, _VP, _U64, C.c_double, C.c_bool, _U64)),
    "numba_r2r_separable_hartley": (None, (_U64, _VP, _VP, _VP, C.c_double, _U64)),
    "numba_r2r_genuine_hartley": (None, (_U64, _VP, _VP, _VP, C.c_double, _U64)),
    "numba_r2r_fftpack": (None, (_U64, _VP, _VP, _VP, C.c_bool, C.c_bool, C.c_double, _U64)),
}
NUMBA_SYMBOLS = tuple(_SIGS)


class LowLevelLib:
    """The ten ``numba_*`` entry points of one shared library, callable on arrays."""

    def __init__(self, path: str):
        self.path = str(path)
        self.cdll = C.CDLL(self.path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.cdll, name)
            fn.restype = res
            fn.argtypes = list(args)

    # --- helpers -----------------------------------------------------------
    def _call(self, name, ain, aout, axes, *rest):
        ndim = len(describe(ain)[1])
        if len(describe(aout)[1]) != ndim:
            raise ValueError("Input and output array must have the same number of dimensions")
        rin, k1 = make_record(ain)
        if aout is ain:
            rout, k2 = rin, k1
        else:
            rout, k2 = make_record(aout)
        rax, k3 = axes_record(axes)
        getattr(self.cdll, name)(ndim, C.addressof(rin), C.addressof(rout), C.addressof(rax), *rest)
        del k1, k2, k3
        return aout

    # --- API (argument order of rocket_fft/__init__.pyi:6-105) --------------
    def good_size(self, n, real):
        return int(self.cdll.numba_good_size(int(n), bool(real)))

    def c2c(self, ain, aout, axes, forward, fct, nthreads=1):
        return self._call("numba_c2c", ain, aout, axes, bool(forward), float(fct), int(nthreads))

    def r2c(self, ain, aout, axes, forward, fct, nthreads=1):
        return self._call("numba_r2c", ain, aout, axes, bool(forward), float(fct), int(nthreads))

    def c2r(self, ain, aout, axes, forward, fct, nthreads=1):
        return self._call("numba_c2r", ain, aout, axes, bool(forward), float(fct), int(nthreads))

    def c2c_sym(self, ain, aout, axes, forward, fct, nthreads=1):
        return self._call("numba_c2c_sym", ain, aout, axes, bool(forward), float(fct), int(nthreads))

    def dct(self, ain, aout, axes, type, fct, ortho, nthreads=1):
        return self._call("numba_dct", ain, aout, axes, int(type), float(fct), bool(ortho), int(nthreads))

    def dst(self, ain, aout, axes, type, fct, ortho, nthreads=1):
        return self._call("numba_dst", ain, aout, axes, int(type), float(fct), bool(ortho), int(nthreads))

    def r2r_separable_hartley(self, ain, aout, axes, fct, nthreads=1):
        return self._call("numba_r2r_separable_hartley", ain, aout, axes, float(fct), int(nthreads))

    def r2r_genuine_hartley(self, ain, aout, axes, fct, nthreads=1):
        return self._call("numba_r2r_genuine_hartley", ain, aout, axes, float(fct), int(nthreads))

    def r2r_fftpack(self, ain, aout, axes, real2hermitian, forward, fct, nthreads=1):
        return self._call(
            "numba_r2r_fftpack", ain, aout, axes, bool(real2hermitian), bool(forward), float(fct), int(nthreads)
        )
