"""Build the drop-in demonstration tree baseline/_ref/ (git-ignored, travels with gpurun):

    baseline/_ref/rocket_fft/   the reference's own Python package, UNMODIFIED (copied at build time
                                from /root/reference -- never committed), with
        _special_helpers<EXT>   compiled from the reference's source (FFTLog helpers, not on the hot path)
        _pocketfft_numba<EXT>   = rocket_fft_b200/librocketfft_b200.so  <-- the swap
    baseline/_ref/tests/        the reference's own test files

The reference locates its extension by file name (rocket_fft/extutils.py:12-18) and binds the ten
numba_* symbols by name (rocket_fft/pocketfft.py:33-128), so with our library under that name its whole
np.fft / scipy.fft overload layer runs on the B200 kernels.  tests/test_gpu_dropin_reference_suite.py
runs the reference's test files against this tree.  Run in the build container:
    python tools/make_dropin_demo.py
"""
import os
import shutil
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
EXT = sysconfig.get_config_var("EXT_SUFFIX")


def main():
    if not os.path.isdir(REF):
        print("no /root/reference here: nothing to do")
        return 0
    pkg = os.path.join(DST, "rocket_fft")
    tst = os.path.join(DST, "tests")
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(pkg)
    os.makedirs(tst)
    for f in os.listdir(os.path.join(REF, "rocket_fft")):
        if f.endswith((".py", ".pyi", ".typed")):
            shutil.copy(os.path.join(REF, "rocket_fft", f), pkg)
    for f in os.listdir(os.path.join(REF, "tests")):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF, "tests", f), tst)
    os.makedirs(os.path.join(tst, "__pycache__"), exist_ok=True)  # the reference's cache-cleanup fixture lists it
    open(os.path.join(tst, "__pycache__", ".keep"), "w").close()
    with open(os.path.join(tst, "conftest.py"), "w") as fh:
        fh.write("import rocket_fft  # noqa: F401  (registers the overloads; the entry point only exists when pip-installed)\n")
    inc = sysconfig.get_paths()["include"]
    subprocess.run(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", f"-I{inc}",
                    os.path.join(REF, "rocket_fft", "_special_helpers.cpp"), "-o",
                    os.path.join(pkg, "_special_helpers" + EXT)], check=True)
    shutil.copy(os.path.join(ROOT, "rocket_fft_b200", "librocketfft_b200.so"), os.path.join(pkg, "_pocketfft_numba" + EXT))
    print("drop-in tree ready:", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
