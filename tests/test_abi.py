"""CPU tests of the boundary: the shared library loads, exports every symbol the
header declares, and the integer-only entry point (good_size) is bit-exact.
No compute entry point is called here (there is no GPU on the CPU runner)."""
import ctypes
import os
import re

import numpy as np
import pytest

import parity

ROOT = parity.ROOT
HEADER = os.path.join(ROOT, "include", "rocketfft_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(numba_\w+|rfb200_\w+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_ten_dropin_symbols():
    from rocket_fft_b200._abi import NUMBA_SYMBOLS

    decl = _declared_symbols()
    for s in NUMBA_SYMBOLS:
        assert s in decl
    assert len([d for d in decl if d.startswith("numba_")]) == 10


def test_library_exports_every_declared_symbol():
    import rocket_fft_b200 as r

    cdll = ctypes.CDLL(r.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(cdll, name), name


def test_good_size_bit_exact_vs_golden_and_reference():
    import rocket_fft_b200 as r

    d, _ = parity.golden()
    for t, gc, gr in zip(d["good_size_targets"], d["good_size_cmplx"], d["good_size_real"]):
        assert r.good_size(int(t), False) == int(gc)
        assert r.good_size(int(t), True) == int(gr)
    ref = parity.reflib()
    if ref is not None:
        rng = np.random.default_rng(0)
        targets = list(range(0, 100000, 7)) + [int(x) for x in rng.integers(1, 2**40, 300)]
        for t in targets:
            assert r.good_size(t, False) == ref.good_size(t, False)
            assert r.good_size(t, True) == ref.good_size(t, True)


def test_argument_validation_is_host_side():
    import rocket_fft_b200 as r

    x = np.zeros(8, dtype=np.complex128)
    with pytest.raises(TypeError):
        r.c2c(x, np.zeros(8, dtype=np.complex64), [0], True, 1.0)
    with pytest.raises(TypeError):
        r.r2c(x, x, [0], True, 1.0)
    with pytest.raises(ValueError):
        r.dct(np.zeros(8), np.zeros(8), [0], 5, 1.0, False)
