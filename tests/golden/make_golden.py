"""Generate tests/golden/ref_vectors.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):
    make -C oracle            # builds oracle/_ref/libpocketfft_ref.so from /root/reference
    python tests/golden/make_golden.py

Every case calls one ``numba_*`` symbol of the compiled reference through ctypes
(rocket_fft_b200._abi.LowLevelLib -- the same record layout Numba passes) on a seeded
input and stores input, arguments and output.  The file also embeds a subset of the
FFTW-generated DCT/DST known-answer vectors SciPy ships
(scipy/fftpack/tests/fftw_double_ref.npz) that the reference's own test-suite uses
(tests/test_scipy_testsuite.py:1192-1293), and the README example (README.md:26-35).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from rocket_fft_b200._abi import LowLevelLib  # noqa: E402

ref = LowLevelLib(os.path.join(ROOT, "oracle", "_ref", "libpocketfft_ref.so"))
rng = np.random.default_rng(20261017)

cases = []  # list of dict(meta) ; arrays stored under f"c{idx}_in" / f"c{idx}_out"
arrays = {}


def cplx(shape, dt):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)


def real(shape, dt):
    return rng.standard_normal(shape).astype(dt)


def add(op, ain, out_shape, out_dtype, axes, **kw):
    aout = np.zeros(out_shape, dtype=out_dtype)
    fn = getattr(ref, op)
    if op in ("c2c", "r2c", "c2r", "c2c_sym"):
        fn(ain, aout, axes, kw["forward"], kw["fct"], 1)
    elif op in ("dct", "dst"):
        fn(ain, aout, axes, kw["type"], kw["fct"], kw["ortho"], 1)
    elif op in ("r2r_separable_hartley", "r2r_genuine_hartley"):
        fn(ain, aout, axes, kw["fct"], 1)
    elif op == "r2r_fftpack":
        fn(ain, aout, axes, kw["real2hermitian"], kw["forward"], kw["fct"], 1)
    idx = len(cases)
    arrays[f"c{idx}_in"] = ain
    arrays[f"c{idx}_out"] = aout
    cases.append(dict(op=op, axes=[int(a) for a in axes], **kw))


# ---- c2c ---------------------------------------------------------------------
for n in (1, 2, 3, 4, 5, 7, 8, 11, 13, 16, 30, 36, 49, 64, 127, 128, 129, 210, 1000, 1024, 2011, 4096):
    for dt in (np.complex128, np.complex64):
        for fwd in (True, False):
            if n > 300 and not fwd:
                continue
            x = cplx((n,), dt)
            add("c2c", x, x.shape, dt, [0], forward=fwd, fct=1.0 if fwd else 1.0 / n)
for shp, axes in (((8, 17, 13), [0, 1, 2]), ((8, 17, 13), [2, 0]), ((16, 9), [0]), ((16, 9), [1, 1]), ((1, 29), [0, 1])):
    for dt in (np.complex128, np.complex64):
        x = cplx(shp, dt)
        add("c2c", x, shp, dt, axes, forward=True, fct=0.5)

# ---- r2c / c2r / c2c_sym -------------------------------------------------------
for shp, axes in (((16,), [0]), ((17,), [0]), ((30,), [0]), ((1,), [0]), ((2,), [0]), ((12, 10), [0, 1]), ((12, 11), [1, 0]),
                  ((6, 7, 8), [0, 1, 2]), ((6, 7, 9), [1]), ((128,), [0]), ((2011,), [0])):
    for dt, cdt in ((np.float64, np.complex128), (np.float32, np.complex64)):
        for fwd in (True, False):
            x = real(shp, dt)
            oshp = list(shp)
            oshp[axes[-1]] = shp[axes[-1]] // 2 + 1
            add("r2c", x, tuple(oshp), cdt, axes, forward=fwd, fct=1.25)
            X = cplx(tuple(oshp), cdt)
            add("c2r", X, shp, dt, axes, forward=fwd, fct=0.75)
            add("c2c_sym", x, shp, cdt, axes, forward=fwd, fct=1.0)

# ---- DCT / DST -----------------------------------------------------------------
for N in (1, 2, 3, 4, 8, 9, 16, 17, 30, 31, 100, 101):
    for op in ("dct", "dst"):
        for t in (1, 2, 3, 4):
            if op == "dct" and t == 1 and N < 2:
                continue
            for ortho in (False, True):
                for dt in (np.float64, np.float32):
                    if dt is np.float32 and N not in (8, 17, 100):
                        continue
                    x = real((N,), dt)
                    add(op, x, x.shape, dt, [0], type=t, fct=1.0, ortho=ortho)
for op in ("dct", "dst"):
    for t in (1, 2, 3, 4):
        x = real((6, 10, 5), np.float64)
        add(op, x, x.shape, np.float64, [0, 1], type=t, fct=0.125, ortho=False)

# ---- Hartley / fftpack -----------------------------------------------------------
for shp, axes in (((10,), [0]), ((127,), [0]), ((16, 12), [0, 1]), ((1, 29), [0, 1]), ((29, 1), [1, 0]), ((8, 5, 9), [0, 1, 2]),
                  ((8, 5, 9), [2, 0])):
    for dt in (np.float64, np.float32):
        x = real(shp, dt)
        add("r2r_separable_hartley", x, shp, dt, axes, fct=1.0)
        add("r2r_genuine_hartley", x, shp, dt, axes, fct=0.5)
for shp, axes in (((3,), [0]), ((8,), [0]), ((9,), [0]), ((10, 7), [0, 1]), ((10, 7), [1])):
    for r2h in (True, False):
        for fwd in (True, False):
            for dt in (np.float64, np.float32):
                x = real(shp, dt)
                add("r2r_fftpack", x, shp, dt, axes, real2hermitian=r2h, forward=fwd, fct=1.0)

# ---- good_size -------------------------------------------------------------------
gs_targets = np.array(list(range(0, 2049)) + [15015, 1000003, 2000005, 999983, 2**31 - 1, 2**31 + 1, 2**40 + 1], dtype=np.uint64)
arrays["good_size_targets"] = gs_targets
arrays["good_size_cmplx"] = np.array([ref.good_size(int(t), False) for t in gs_targets], dtype=np.uint64)
arrays["good_size_real"] = np.array([ref.good_size(int(t), True) for t in gs_targets], dtype=np.uint64)

# ---- README example (README.md:26-35) ----------------------------------------------
arrays["readme_in"] = np.array([1, 6, 1, 8, 0, 3, 3, 9], dtype=np.complex128)
o = np.empty(8, dtype=np.complex128)
ref.c2c(arrays["readme_in"], o, [0], True, 1.0, 1)
arrays["readme_out"] = o

# ---- FFTW known answers shipped by SciPy -------------------------------------------
try:
    import scipy.fftpack

    d = os.path.join(os.path.dirname(scipy.fftpack.__file__), "tests", "fftw_double_ref.npz")
    fw = np.load(d)
    for N in (2, 3, 4, 8, 12, 15, 16, 17, 32, 64):
        for kind in ("dct", "dst"):
            for t in (1, 2, 3, 4):
                key = f"{kind}_{t}_{N}"
                if key in fw.files:
                    arrays["fftw_" + key] = fw[key]
except Exception as e:  # pragma: no cover
    print("FFTW vectors not embedded:", e)

arrays["cases_json"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
out = os.path.join(HERE, "ref_vectors.npz")
np.savez_compressed(out, **arrays)
print(len(cases), "cases ->", out, os.path.getsize(out), "bytes")
