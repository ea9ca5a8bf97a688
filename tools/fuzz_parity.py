"""Randomised parity fuzz against the compiled reference (oracle/_ref): random ops, dtypes, shapes,
strides (sliced / reversed / transposed views, non-contiguous outputs), axes (incl. repeats for the ops
where the reference defines them), flags.  Usage: python tools/fuzz_parity.py [seconds] [seed]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import parity
import rocket_fft_b200 as R

T = parity.reflib()
assert T is not None, "oracle/_ref/libpocketfft_ref.so missing"
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
LENS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 17, 20, 24, 25, 27, 30, 31, 32, 33, 36, 45, 48, 49, 60, 64, 67, 77, 81,
        96, 100, 101, 120, 121, 125, 127, 128, 129, 143, 169, 200, 243, 256, 257, 300, 343, 360, 384, 500, 512, 625, 729, 1000, 1024]


def view(a):
    """random strided view of a freshly allocated larger array with the same content"""
    kind = rng.integers(0, 5)
    if kind == 0:
        return a
    if kind == 1:
        return np.asfortranarray(a)
    if kind == 2:  # reversed along a random axis
        ax = rng.integers(0, a.ndim)
        big = np.ascontiguousarray(np.flip(a, ax))
        return np.flip(big, ax)
    if kind == 3:  # every other element of a padded array along a random axis
        ax = rng.integers(0, a.ndim)
        shp = list(a.shape)
        shp[ax] *= 2
        big = np.zeros(shp, dtype=a.dtype)
        sl = [slice(None)] * a.ndim
        sl[ax] = slice(0, None, 2)
        big[tuple(sl)] = a
        return big[tuple(sl)]
    perm = rng.permutation(a.ndim)
    big = np.ascontiguousarray(a.transpose(perm))
    return big.transpose(np.argsort(perm))


def empty_like_view(shape, dt):
    return view(np.zeros(shape, dtype=dt))


def rand(shape, dt):
    if np.dtype(dt).kind == "c":
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)
    return rng.standard_normal(shape).astype(dt)


t0 = time.time()
ncase = 0
worst = 0.0
while time.time() - t0 < budget:
    nd = int(rng.integers(1, 4))
    shape = tuple(int(rng.choice(LENS[: 40 if nd > 1 else len(LENS)])) for _ in range(nd))
    while np.prod(shape) > 400000:
        shape = tuple(max(1, s // 2) for s in shape)
    single = bool(rng.integers(0, 2))
    rdt, cdt = (np.float32, np.complex64) if single else (np.float64, np.complex128)
    op = rng.choice(["c2c", "r2c", "c2r", "c2c_sym", "dct", "dst", "sep", "gen", "pack"])
    nax = int(rng.integers(1, nd + 1))
    axes = [int(a) for a in rng.permutation(nd)[:nax]]
    if op in ("c2c", "dct", "dst", "sep", "pack") and rng.integers(0, 6) == 0:
        axes.append(axes[0])  # repeated axis (numpy-like mode of the reference)
    fwd = bool(rng.integers(0, 2))
    fct = float(rng.choice([1.0, 0.5, 1.0 / max(1, shape[axes[-1]])]))
    n = int(np.prod([shape[a] for a in set(axes)]))
    what = (op, shape, axes, single, fwd, fct)
    try:
        if op == "c2c":
            x = view(rand(shape, cdt))
            a, b = empty_like_view(shape, cdt), np.empty(shape, dtype=cdt)
            R.c2c(x, a, axes, fwd, fct); T.c2c(x, b, axes, fwd, fct)
        elif op in ("r2c", "c2c_sym"):
            x = view(rand(shape, rdt))
            oshp = list(shape)
            if op == "r2c":
                oshp[axes[-1]] = shape[axes[-1]] // 2 + 1
            a, b = empty_like_view(tuple(oshp), cdt), np.zeros(oshp, dtype=cdt)
            getattr(R, op)(x, a, axes, fwd, fct); getattr(T, op)(x, b, axes, fwd, fct)
        elif op == "c2r":
            ishp = list(shape)
            ishp[axes[-1]] = shape[axes[-1]] // 2 + 1
            x = view(rand(tuple(ishp), cdt))
            a, b = empty_like_view(shape, rdt), np.empty(shape, dtype=rdt)
            R.c2r(x, a, axes, fwd, fct); T.c2r(x, b, axes, fwd, fct)
        elif op in ("dct", "dst"):
            t = int(rng.integers(1, 5))
            ortho = bool(rng.integers(0, 2))
            if op == "dct" and t == 1 and min(shape[a] for a in axes) < 2:
                continue
            what += (t, ortho)
            x = view(rand(shape, rdt))
            a, b = empty_like_view(shape, rdt), np.empty(shape, dtype=rdt)
            getattr(R, op)(x, a, axes, t, fct, ortho); getattr(T, op)(x, b, axes, t, fct, ortho)
            n = n * 4
        elif op in ("sep", "gen"):
            x = view(rand(shape, rdt))
            a, b = empty_like_view(shape, rdt), np.empty(shape, dtype=rdt)
            f = R.r2r_separable_hartley if op == "sep" else R.r2r_genuine_hartley
            g = T.r2r_separable_hartley if op == "sep" else T.r2r_genuine_hartley
            f(x, a, axes, fct); g(x, b, axes, fct)
        else:
            r2h = bool(rng.integers(0, 2))
            what += (r2h,)
            x = view(rand(shape, rdt))
            a, b = empty_like_view(shape, rdt), np.empty(shape, dtype=rdt)
            R.r2r_fftpack(x, a, axes, r2h, fwd, fct); T.r2r_fftpack(x, b, axes, r2h, fwd, fct)
    except Exception as e:
        print("EXCEPTION", what, repr(e))
        sys.exit(2)
    err = parity.l2err(np.asarray(a), b)
    tol = parity.tol(rdt, max(n, 2))
    worst = max(worst, err / tol)
    ncase += 1
    if not (err <= tol):
        print("MISMATCH", what, "err", err, "tol", tol, "strides in", x.strides, "out", a.strides)
        sys.exit(1)
print(f"fuzz ok: {ncase} random cases in {time.time() - t0:.0f} s, worst err/tol = {worst:.3f}, launches {R.launch_count()}")
