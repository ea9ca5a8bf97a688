"""Generate tests/golden/fht_vectors.npz from the UNMODIFIED reference (its scipy.fft.fht / ifht / fhtoffset
Numba overloads, rocket_fft/overloads.py:1755-1859).

Run in the build container (where /root/reference exists), after `make -C oracle`:
    python tests/golden/make_fht_golden.py

A throw-away tree under /tmp holds the reference's Python package next to its two native helpers compiled from
the reference's own sources (the transform library = oracle/_ref/libpocketfft_ref.so under the extension's file
name, and _special_helpers); nothing of it is copied into this repository.
"""
import json
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
EXT = sysconfig.get_config_var("EXT_SUFFIX")

tree = tempfile.mkdtemp(prefix="rocketfft_ref_")
pkg = os.path.join(tree, "rocket_fft")
os.makedirs(pkg)
for f in os.listdir(os.path.join(REF, "rocket_fft")):
    if f.endswith((".py", ".pyi", ".typed")):
        shutil.copy(os.path.join(REF, "rocket_fft", f), pkg)
shutil.copy(os.path.join(ROOT, "oracle", "_ref", "libpocketfft_ref.so"), os.path.join(pkg, "_pocketfft_numba" + EXT))
subprocess.run(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-I" + sysconfig.get_paths()["include"],
                os.path.join(REF, "rocket_fft", "_special_helpers.cpp"), "-o", os.path.join(pkg, "_special_helpers" + EXT)],
               check=True)
sys.path.insert(0, tree)

import numba as nb  # noqa: E402
import scipy.fft  # noqa: E402

import rocket_fft  # noqa: E402,F401  (registers the overloads)


@nb.njit
def ref_fht(a, dln, mu, offset, bias):
    return scipy.fft.fht(a, dln, mu, offset, bias)


@nb.njit
def ref_ifht(a, dln, mu, offset, bias):
    return scipy.fft.ifht(a, dln, mu, offset, bias)


@nb.njit
def ref_fhtoffset(dln, mu, initial, bias):
    return scipy.fft.fhtoffset(dln, mu, initial, bias)


rng = np.random.default_rng(20261018)
cases, arrays = [], {}
for dt in ("float64", "float32"):
    for shape in ((64,), (3, 128), (2, 5, 33), (4, 1000), (1,), (2, 2)):
        for dln, mu, offset, bias in ((0.1, 0.0, 0.0, 0.0), (0.05, 1.5, 0.3, 0.0), (0.2, 0.5, -0.4, 0.3), (0.1, 2.0, 0.1, -0.6),
                                      (0.1, -0.5, 0.0, 0.0)):
            a = rng.standard_normal(shape).astype(dt)
            i = len(cases)
            arrays[f"c{i}_in"] = a
            arrays[f"c{i}_fht"] = ref_fht(a, dln, mu, offset, bias)
            arrays[f"c{i}_ifht"] = ref_ifht(a, dln, mu, offset, bias)
            cases.append(dict(dtype=dt, shape=list(shape), dln=dln, mu=mu, offset=offset, bias=bias))
offs = []
for dln, mu, initial, bias in ((0.1, 0.0, 0.0, 0.0), (0.05, 1.5, 0.3, 0.0), (0.2, 0.5, -0.4, 0.3), (0.01, 3.0, 2.0, -0.5)):
    offs.append(dict(dln=dln, mu=mu, initial=initial, bias=bias, value=float(ref_fhtoffset(dln, mu, initial, bias))))
arrays["cases_json"] = np.frombuffer(json.dumps(dict(cases=cases, offsets=offs)).encode(), dtype=np.uint8)
out = os.path.join(HERE, "fht_vectors.npz")
np.savez_compressed(out, **arrays)
shutil.rmtree(tree, ignore_errors=True)
print("wrote", out, os.path.getsize(out), "bytes,", len(cases), "cases")
