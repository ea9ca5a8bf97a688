"""Numba bindings: the low-level functions usable inside ``@njit`` code.

Same names, argument order and typing rules as the reference's low-level interface
(rocket_fft/pocketfft.py:147-226; README.md:74-85), re-implemented here: each function is a
Numba intrinsic that passes pointers to the arrays' native records to the ``numba_*`` symbol
of ``librocketfft_b200.so`` -- the drop-in replacement of the symbol the reference's
intrinsic calls (rocket_fft/pocketfft.py:33-128).  The library is made visible to the JIT
linker with ``llvmlite.binding.load_library_permanently`` exactly as the reference does for
its own extension (rocket_fft/extutils.py:20-21).

Host (NumPy) arrays are staged to the GPU and back inside the call; there is no CPU
transform path.  Typing rules (checked at compile time, mirroring the reference's tests
tests/test_low_level_interface.py:200-325): `ain`/`aout` arrays of equal ndim, `axes` a 1-D
array (any integer/float dtype; converted to uint64 when needed), `forward`/`ortho`/
`real2hermitian` booleans, `fct` a float, `nthreads`/`type` integers.
"""
from __future__ import annotations

from llvmlite import binding as _llb
from llvmlite import ir as _ir
from numba.core import cgutils as _cg
from numba.core import types as _nt
from numba.core.errors import TypingError
from numba.extending import intrinsic as _intrinsic
from numba.np import arrayobj as _arrayobj

from .lowlevel import LIB_PATH

_llb.load_library_permanently(LIB_PATH)

_I64 = _ir.IntType(64)
_I1 = _ir.IntType(1)
_F64 = _ir.DoubleType()
_PTR = _ir.IntType(8).as_pointer()

# kinds of scalar arguments after (ain, aout, axes)
_BOOL, _FLOAT, _INT = "bool", "float", "int"
_LL = {_BOOL: _I1, _FLOAT: _F64, _INT: _I64}


def _check_scalar(name, ty, kind):
    if kind == _BOOL and not isinstance(ty, _nt.Boolean):
        raise TypingError(f"'{name}' must be a boolean")
    if kind == _FLOAT and not isinstance(ty, _nt.Float):
        raise TypingError(f"'{name}' must be a float")
    if kind == _INT and not isinstance(ty, _nt.Integer):
        raise TypingError(f"'{name}' must be an integer")


def _record_ptr(context, builder, ary_ty, ary_val):
    """i8* to the native array record {meminfo, parent, nitems, itemsize, data, shape, strides}."""
    rec = _arrayobj.make_array(ary_ty)(context, builder, ary_val)
    return builder.bitcast(rec._getpointer(), _PTR)


def _widen(builder, val, kind):
    if kind == _INT and val.type.width != 64:
        return builder.zext(val, _I64)
    if kind == _FLOAT and val.type != _F64:
        return builder.fpext(val, _F64)
    return val


def _binding(symbol, scalars):
    """Build the intrinsic for `symbol(ndim, ain*, aout*, axes*, <scalars...>) -> void`."""
    names = [n for n, _ in scalars]
    kinds = [k for _, k in scalars]

    def typer_and_codegen(typingctx, ain, aout, axes, *rest):
        if not (isinstance(ain, _nt.Array) and isinstance(aout, _nt.Array) and isinstance(axes, _nt.Array)):
            raise TypingError("ain, aout and axes must be arrays")
        if ain.ndim != aout.ndim:
            raise TypingError("Input and output array must have the same number of dimensions")
        if axes.ndim != 1:
            raise TypingError("Axes must be a one-dimensional array")
        for nm, ty, kd in zip(names, rest, kinds):
            _check_scalar(nm, ty, kd)
        native_axes = isinstance(axes.dtype, _nt.Integer) and axes.dtype.bitwidth == 64 and axes.layout in "CF"
        sig = _nt.void(ain, aout, axes, *rest)

        def codegen(context, builder, signature, args):
            a_in, a_out, a_ax = args[:3]
            ax_ty = signature.args[2]
            if not native_axes:
                u64_1d = _nt.Array(_nt.uint64, 1, "C")
                a_ax = _arrayobj.array_astype(context, builder, u64_1d(ax_ty, _nt.uint64), (a_ax, _nt.uint64))
                ax_ty = u64_1d
            call_args = [
                _I64(signature.args[0].ndim),
                _record_ptr(context, builder, signature.args[0], a_in),
                _record_ptr(context, builder, signature.args[1], a_out),
                _record_ptr(context, builder, ax_ty, a_ax),
            ]
            for val, kd in zip(args[3:], kinds):
                call_args.append(_widen(builder, val, kd))
            fnty = _ir.FunctionType(_ir.VoidType(), [_I64, _PTR, _PTR, _PTR] + [_LL[k] for k in kinds])
            fn = _cg.get_or_insert_function(builder.module, fnty, symbol)
            builder.call(fn, call_args)
            return context.get_dummy_value()

        return sig, codegen

    # numba's @intrinsic inspects the signature: build a function with explicit parameter names
    params = ", ".join(names)
    src = (
        f"def {symbol[6:]}(typingctx, ain, aout, axes, {params}):\n"
        f"    return _impl(typingctx, ain, aout, axes, {params})\n"
    )
    ns = {"_impl": typer_and_codegen}
    exec(src, ns)
    return _intrinsic(ns[symbol[6:]])


_cmplx = [("forward", _BOOL), ("fct", _FLOAT), ("nthreads", _INT)]
_real = [("type", _INT), ("fct", _FLOAT), ("ortho", _BOOL), ("nthreads", _INT)]
_hart = [("fct", _FLOAT), ("nthreads", _INT)]
_pack = [("real2hermitian", _BOOL), ("forward", _BOOL), ("fct", _FLOAT), ("nthreads", _INT)]

c2c = _binding("numba_c2c", _cmplx)
r2c = _binding("numba_r2c", _cmplx)
c2r = _binding("numba_c2r", _cmplx)
c2c_sym = _binding("numba_c2c_sym", _cmplx)
dct = _binding("numba_dct", _real)
dst = _binding("numba_dst", _real)
r2r_separable_hartley = _binding("numba_r2r_separable_hartley", _hart)
r2r_genuine_hartley = _binding("numba_r2r_genuine_hartley", _hart)
r2r_fftpack = _binding("numba_r2r_fftpack", _pack)
separable_hartley = r2r_separable_hartley
genuine_hartley = r2r_genuine_hartley
fftpack = r2r_fftpack


@_intrinsic
def good_size(typingctx, n, real):
    if not isinstance(n, (_nt.Integer, _nt.Boolean)):
        raise TypingError("The first argument 'n' must be an integer")
    if not isinstance(real, (_nt.Integer, _nt.Boolean)):
        raise TypingError("The second argument 'real' must be a boolean")

    def codegen(context, builder, signature, args):
        n_v, real_v = args
        if n_v.type.width != 64:
            n_v = builder.zext(n_v, _I64)
        if real_v.type.width != 1:
            real_v = builder.trunc(real_v, _I1)  # as the reference does (pocketfft.py:220)
        fnty = _ir.FunctionType(_I64, [_I64, _I1])
        fn = _cg.get_or_insert_function(builder.module, fnty, "numba_good_size")
        return builder.call(fn, [n_v, real_v])

    return _nt.uint64(n, real), codegen
