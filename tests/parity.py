"""Shared helpers for the parity tests: the reference library (if built), golden
vectors, the error metric and the north-star tolerances."""
import json
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpocketfft_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")

_ref = None


def reflib():
    """The compiled, unmodified reference (oracle/_ref), or None when it is absent."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        from oracle.abi_view import RefLib  # neutral module: never maps the product's .so

        _ref = RefLib(REF_SO)
    return _ref


def golden():
    d = np.load(GOLDEN)
    cases = json.loads(bytes(d["cases_json"]).decode())
    return d, cases


def l2err(a, b):
    """rel-L2 = ||a-b|| / ||b||  (same definition as the reference's
    tests/test_low_level_interface.py:72-73); b is the trusted side."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.sqrt(np.sum(np.abs(b.astype(np.complex128)) ** 2))
    num = np.sqrt(np.sum(np.abs(a.astype(np.complex128) - b.astype(np.complex128)) ** 2))
    if den == 0:
        return float(num)
    return float(num / den)


def tol(dtype, n):
    """North-star bound: rel-L2 <= 1e-5*log2(n) for fp32, 1e-13*log2(n) for fp64,
    n = transformed length (product over the transformed axes)."""
    lg = max(1.0, math.log2(max(int(n), 2)))
    single = np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
    return (1e-5 if single else 1e-13) * lg


def call(lib, case, ain, aout):
    """Invoke one golden-style case dict on a LowLevelLib-like object."""
    op = case["op"]
    ax = case["axes"]
    if op in ("c2c", "r2c", "c2r", "c2c_sym"):
        return getattr(lib, op)(ain, aout, ax, case["forward"], case["fct"], 1)
    if op in ("dct", "dst"):
        return getattr(lib, op)(ain, aout, ax, case["type"], case["fct"], case["ortho"], 1)
    if op in ("r2r_separable_hartley", "r2r_genuine_hartley"):
        return getattr(lib, op)(ain, aout, ax, case["fct"], 1)
    if op == "r2r_fftpack":
        return lib.r2r_fftpack(ain, aout, ax, case["real2hermitian"], case["forward"], case["fct"], 1)
    raise ValueError(op)


def tlen(case, shape):
    n = 1
    for a in set(case["axes"]):
        n *= shape[a]
    return n
