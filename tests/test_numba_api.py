"""CPU tests of the Numba bindings: typing rules and symbol resolution (compile only --
nothing is executed on the transform path, which needs a GPU); good_size runs (host code)."""
import numba as nb
import numpy as np
import pytest

from rocket_fft_b200 import numba_api as napi


def test_good_size_inside_njit():
    @nb.njit
    def gs(n, real):
        return napi.good_size(n, real)

    assert gs(1000003, False) == 1000188
    assert gs(1000003, True) == 1012500
    assert gs(np.int32(15015), np.bool_(False)) == 15092
    with pytest.raises(Exception):
        gs(1.0, True)


def test_low_level_functions_compile_and_link():
    sigs = {
        "c2c": ("void(c16[:,:], c16[:,:], i8[:], b1, f8, i8)", lambda a, b, ax, f, s, t: napi.c2c(a, b, ax, f, s, t)),
        "r2c": ("void(f4[:], c8[:], u8[:], b1, f8, i8)", lambda a, b, ax, f, s, t: napi.r2c(a, b, ax, f, s, t)),
        "c2r": ("void(c8[:], f4[:], i4[:], b1, f4, i2)", lambda a, b, ax, f, s, t: napi.c2r(a, b, ax, f, s, t)),
        "c2c_sym": ("void(f8[:], c16[:], f8[:], b1, f8, u1)", lambda a, b, ax, f, s, t: napi.c2c_sym(a, b, ax, f, s, t)),
        "dct": ("void(f8[:,:], f8[:,:], i8[:], i8, f8, b1, i8)", lambda a, b, ax, ty, s, o, t: napi.dct(a, b, ax, ty, s, o, t)),
        "dst": ("void(f4[:,:], f4[:,:], i8[:], i8, f8, b1, i8)", lambda a, b, ax, ty, s, o, t: napi.dst(a, b, ax, ty, s, o, t)),
        "sep": ("void(f8[:], f8[:], i8[:], f8, i8)", lambda a, b, ax, s, t: napi.r2r_separable_hartley(a, b, ax, s, t)),
        "gen": ("void(f8[:], f8[:], i8[:], f8, i8)", lambda a, b, ax, s, t: napi.r2r_genuine_hartley(a, b, ax, s, t)),
        "pack": ("void(f8[:], f8[:], i8[:], b1, b1, f8, i8)", lambda a, b, ax, r, f, s, t: napi.r2r_fftpack(a, b, ax, r, f, s, t)),
    }
    for name, (sig, fn) in sigs.items():
        compiled = nb.njit(sig, nogil=True)(fn)
        assert compiled.signatures, name
        ir = compiled.inspect_llvm(compiled.signatures[0])
        assert "@numba_" in ir, name


def test_typing_errors():
    a = np.zeros(8, dtype=np.complex128)
    ax = np.array([0], dtype=np.uint64)

    @nb.njit
    def f(a, b, ax, fw, fct, nt):
        napi.c2c(a, b, ax, fw, fct, nt)

    for bad in (
        (a, a, ax.reshape(1, 1), True, 1.0, 1),   # axes not 1-D
        (a, a, ax, 1, 1.0, 1),                    # forward not boolean
        (a, a, ax, True, np.int64(1), 1),         # fct not float
        (a, a, ax, True, 1.0, 1.0),               # nthreads not integer
        (a.reshape(2, 4), a, ax, True, 1.0, 1),   # ndim mismatch
    ):
        with pytest.raises(Exception):
            f(*bad)
