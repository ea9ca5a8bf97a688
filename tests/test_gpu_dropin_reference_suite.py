"""The drop-in claim, end to end: the reference's OWN Python package (np.fft / scipy.fft overloads,
Numba intrinsics -- unmodified) and the reference's OWN test files, with our CUDA library in the
place of its `_pocketfft_numba` extension (tree built by tools/make_dropin_demo.py into the
git-ignored baseline/_ref/, which travels to the GPU box).  Skipped when that tree is absent."""
import os
import shutil
import subprocess
import sys
import sysconfig

import pytest

import parity

pytestmark = pytest.mark.gpu

DST = os.path.join(parity.ROOT, "baseline", "_ref")
EXT = sysconfig.get_config_var("EXT_SUFFIX")

# The reference's test files run by default (about 6 minutes on the B200 box): the direct C-ABI tests, the np.fft /
# scipy.fft comparison grids (3016 + 172 + 82 + 14 tests), NumPy's own suite as the reference vendors it, threading,
# and the on-disk JIT-cache test (which re-runs the library in fresh processes).  Left out: test_multiprocessing.py -- it
# transforms in the parent and then in a fork()ed Pool, and a CUDA context does not survive fork(): the children get the
# library's "forked after CUDA was used" error (tests/test_gpu_robustness.py pins that behaviour).
# RFB200_FULL_REFERENCE_SUITE=1 adds SciPy's vendored suite, FFTLog, the scipy-like policy and typing files.
FILES = ["test_low_level_interface.py", "test_overwrite_and_dtype.py", "test_numpy_like.py", "test_numpy_compare.py",
         "test_scipy_compare.py", "test_numpy_testsuite.py", "test_misc.py", "test_multithreading.py", "test_caching.py"]
if os.environ.get("RFB200_FULL_REFERENCE_SUITE"):
    FILES += ["test_scipy_testsuite.py", "test_fftlog.py", "test_scipy_like.py", "test_typing.py"]


@pytest.mark.parametrize("fname", FILES)
def test_reference_test_file_passes_on_the_gpu_library(fname):
    pkg = os.path.join(DST, "rocket_fft")
    if not os.path.isdir(pkg) or not os.path.exists(os.path.join(DST, "tests", fname)):
        pytest.skip("baseline/_ref drop-in tree not built (tools/make_dropin_demo.py)")
    import rocket_fft_b200 as R

    shutil.copy(R.LIB_PATH, os.path.join(pkg, "_pocketfft_numba" + EXT))  # always the current build
    os.makedirs(os.path.join(DST, "tests", "__pycache__"), exist_ok=True)  # listed by the reference's cleanup fixture
    env = dict(os.environ, PYTHONPATH=DST + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(DST, "tests", fname), "-q", "-x", "--tb=short", "-p", "no:cacheprovider"],
                       capture_output=True, text=True, env=env, cwd=os.path.join(DST, "tests"), timeout=3000)
    tail = (r.stdout[-1500:] + r.stderr[-500:])
    assert r.returncode == 0, tail
    assert "passed" in r.stdout
