"""rocket_fft_b200 -- B200-native (sm_100a) transform engine behind rocket-fft's
low-level API.  See lowlevel.py for the interface and DESIGN.md for the design."""
from .lowlevel import (  # noqa: F401
    LIB_PATH,
    TransformError,
    c2c,
    c2c_pad,
    c2c_scatter,
    c2c_sym,
    c2r,
    c2r_pad,
    dct,
    dst,
    failure_count,
    fftpack,
    genuine_hartley,
    good_size,
    last_error,
    launch_count,
    launch_count_reset,
    launch_trace,
    launch_trace_get,
    lib,
    pinned,
    plan_cache_clear,
    plan_cache_stats,
    r2c,
    r2c_pad,
    roll,
    scale_lines,
    r2r_fftpack,
    r2r_genuine_hartley,
    r2r_separable_hartley,
    separable_hartley,
    set_dst_ortho_quirk,
    set_stream,
    version,
)

from .fft import get_workers, numpy_like, scipy_like, set_workers  # noqa: F401,E402  (rocket_fft/__init__.py:3-15)

__version__ = "0.1.0"
