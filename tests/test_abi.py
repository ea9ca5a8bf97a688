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
        # every target up to 10^6 and the large probes of SURVEY.md appendix B
        targets = list(range(0, 1000001)) + [2**31 - 1, 2**31 + 1, 2**40 + 1, 1000003, 2000005, 15015]
        targets += [int(x) for x in rng.integers(1, 2**40, 300)]
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


def test_fused_fourstep_ticket_order_never_puts_a_consumer_first():
    """Scheduling invariant of the fused four-step kernel (csrc/pow2_fused4_kernel.cuh), checked on the host through
    rfb200_debug_fuse4_unit: every (step, strip) appears exactly once, B(s) comes after A(s), and A(s) comes after
    B(s - ring) for every ring > lag -- a tile that waits can only wait for tickets drawn earlier (no deadlock)."""
    import ctypes as C

    import rocket_fft_b200 as R

    f = C.CDLL(R.LIB_PATH).rfb200_debug_fuse4_unit
    f.restype = C.c_int64
    f.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    for S in (1, 2, 3, 5, 8, 33, 132):
        for lag in range(1, min(S, 6) + 1):
            pos = {}
            for u in range(2 * S):
                r = f(u, S, lag)
                assert r >= 0
                key = (r >> 32, r & 0xFFFFFFFF)
                assert key not in pos and key[1] < S
                pos[key] = u
            assert f(2 * S, S, lag) == -1
            assert len(pos) == 2 * S
            for s in range(S):
                assert pos[(0, s)] < pos[(1, s)]
                for ring in (lag + 1, lag + 2, lag + 5):
                    if s >= ring:
                        assert pos[(1, s - ring)] < pos[(0, s)], (S, lag, ring, s)
    assert f(0, 2, 3) == -1


def test_transform_without_a_gpu_fails_loudly():
    """There is no CPU fallback: on a machine without a CUDA device a numba_* call must not return silently with an
    untouched output -- the message is readable, the host output is NaN, the failure is counted."""
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import rocket_fft_b200 as R

    before = R.failure_count()
    x = np.ones((4, 16), dtype=np.complex64)
    y = np.zeros((6, 16), dtype=np.complex64)[::2, ::2]  # strided output: only its own elements may be touched
    x2 = x[:3, :8]
    with pytest.raises(R.TransformError):
        R.c2c(x2, y, [1], True, 1.0)
    assert R.failure_count() == before + 1 and R.last_error()
    assert np.isnan(y.real).all() and np.isnan(y.imag).all()
    full = y.base
    assert (full[1::2] == 0).all() and (full[:, 1::2] == 0).all()
    assert R.good_size(1000003, False) == 1000188  # the host-only function still works
