"""ctypes view of the reference's ``numba_*`` C ABI -- TEST INFRASTRUCTURE ONLY.

Builds Numba-layout array records (numba ``_arraystruct.h``; read by the reference at
rocket_fft/_pocketfft_numba.cpp:31-49) from NumPy arrays and calls the ten entry points of a
shared library that exports them -- in practice ``oracle/_ref/libpocketfft_ref.so``, the
unmodified reference compiled by oracle/Makefile.  Host arrays only.

This module is deliberately independent of the product package: importing it does NOT import
``rocket_fft_b200`` and therefore never maps ``librocketfft_b200.so`` into the process, so that
``bench.py --impl reference`` and the parity tests' trusted side run on reference code alone.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_intp = C.c_ssize_t
_REC = {}


def _record_type(ndim: int):
    t = _REC.get(ndim)
    if t is None:
        class Rec(C.Structure):
            _fields_ = [("meminfo", C.c_void_p), ("parent", C.c_void_p), ("nitems", _intp), ("itemsize", _intp),
                        ("data", C.c_void_p), ("shape_and_strides", _intp * (2 * max(ndim, 1)))]

        t = _REC[ndim] = Rec
    return t


def _record(a: np.ndarray):
    if not isinstance(a, np.ndarray):
        raise TypeError("the reference library works on NumPy arrays")
    r = _record_type(a.ndim)()
    r.meminfo = None
    r.parent = None
    r.nitems = a.size
    r.itemsize = a.itemsize
    r.data = a.ctypes.data
    for i in range(a.ndim):
        r.shape_and_strides[i] = a.shape[i]
        r.shape_and_strides[a.ndim + i] = a.strides[i]
    return r


_VP, _U64, _B, _D = C.c_void_p, C.c_uint64, C.c_bool, C.c_double
# reference: rocket_fft/pocketfft.py:33-128 (LLVM signatures of the ten symbols)
_SIGS = {
    "numba_good_size": (_U64, (_U64, _B)),
    "numba_c2c": (None, (_U64, _VP, _VP, _VP, _B, _D, _U64)),
    "numba_r2c": (None, (_U64, _VP, _VP, _VP, _B, _D, _U64)),
    "numba_c2r": (None, (_U64, _VP, _VP, _VP, _B, _D, _U64)),
    "numba_c2c_sym": (None, (_U64, _VP, _VP, _VP, _B, _D, _U64)),
    "numba_dct": (None, (_U64, _VP, _VP, _VP, _U64, _D, _B, _U64)),
    "numba_dst": (None, (_U64, _VP, _VP, _VP, _U64, _D, _B, _U64)),
    "numba_r2r_separable_hartley": (None, (_U64, _VP, _VP, _VP, _D, _U64)),
    "numba_r2r_genuine_hartley": (None, (_U64, _VP, _VP, _VP, _D, _U64)),
    "numba_r2r_fftpack": (None, (_U64, _VP, _VP, _VP, _B, _B, _D, _U64)),
}


class RefLib:
    """The ten ``numba_*`` entry points of one shared library, callable on NumPy arrays
    (argument order of rocket_fft/__init__.pyi:6-105)."""

    def __init__(self, path: str):
        self.path = str(path)
        self.cdll = C.CDLL(self.path)
        for name, (res, args) in _SIGS.items():
            f = getattr(self.cdll, name)
            f.restype = res
            f.argtypes = list(args)

    def _call(self, name, ain, aout, axes, *rest):
        if ain.ndim != aout.ndim:
            raise ValueError("Input and output array must have the same number of dimensions")
        rin = _record(ain)
        rout = rin if aout is ain else _record(aout)
        ax = np.ascontiguousarray(np.asarray(axes).astype(np.uint64, copy=False).ravel())
        rax = _record(ax)
        getattr(self.cdll, name)(ain.ndim, C.addressof(rin), C.addressof(rout), C.addressof(rax), *rest)
        return aout

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
        return self._call("numba_r2r_fftpack", ain, aout, axes, bool(real2hermitian), bool(forward), float(fct),
                          int(nthreads))
