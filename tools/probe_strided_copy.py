"""Bandwidth ceiling of the four-step column passes' access pattern (see csrc/probe.cu): pure copies of 128-row x 256-byte
tiles, rows 128 array rows apart, against the transform passes themselves.  Usage: python tools/probe_strided_copy.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rocket_fft_b200 as R

lib = C.CDLL(R.LIB_PATH)
f = lib.rfb200_debug_tile_copy
f.restype = C.c_int
f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_void_p]
dev = torch.device("cuda:0")
rows, cols = 16384, 8193
X = torch.randn(rows, cols, dtype=torch.complex64, device=dev)
Y = torch.empty_like(X)
D = torch.empty(((cols + 31) // 32) * rows * 32, dtype=torch.complex64, device=dev)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


s = torch.cuda.current_stream().cuda_stream
nbytes = 2 * X.numel() * 8
for mode, name, src, dst in ((0, "strided -> strided", X, Y), (1, "strided -> dense  ", X, D), (2, "dense   -> strided", D, Y)):
    ms = timeit(lambda: f(src.data_ptr(), dst.data_ptr(), rows, cols, cols * 8, mode, s))
    print(f"tile copy {name}: {ms:.4f} ms  {nbytes / ms / 1e6:7.0f} GB/s  {nbytes / ms / 1e6 / 6527.8 * 100:5.1f}% of the HBM copy peak")
ms = timeit(lambda: Y.copy_(X))
print(f"dense copy_                  : {ms:.4f} ms  {nbytes / ms / 1e6:7.0f} GB/s")
ms = timeit(lambda: R.c2c(X, X, [0], True, 1.0))
print(f"c2c columns (two passes)     : {ms:.4f} ms  ({2 * nbytes / ms / 1e6:7.0f} GB/s over both passes)")

# ---- other strided passes of the BASELINE configs, as pure copies (segment bytes x rows per CTA, threads per CTA) ----------
g = lib.rfb200_debug_seg_copy
g.restype = C.c_int
g.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int64, C.c_uint32, C.c_uint32, C.c_int64, C.c_uint32, C.c_uint32, C.c_void_p]
del X, Y, D
torch.cuda.empty_cache()
V = torch.randn(512, 1024, 1024, dtype=torch.complex64, device=dev)  # half of config 3 (4 GiB)
Wv = torch.empty_like(V)
vb = 2 * V.numel() * 8
KiB, MiB = 1024, 1024 * 1024
cases = [
    # name, nrows, seg, row_stride, tiles0, tiles1, outer_stride, threads
    ("cfg3 axis 1: 1024 rows x 128 B, 8 KiB apart (W=16)", 1024, 128, 8 * KiB, 64, 512, 8 * MiB, 1024),
    ("cfg3 axis 1: 1024 rows x  64 B, 8 KiB apart (W=8) ", 1024, 64, 8 * KiB, 128, 512, 8 * MiB, 512),
    ("cfg3 axis 0:  512 rows x 128 B, 8 MiB apart (W=16)", 512, 128, 8 * MiB, 64, 1024, 8 * KiB, 512),
    ("cfg3 axis 0:  512 rows x 256 B, 8 MiB apart (W=32)", 512, 256, 8 * MiB, 32, 1024, 8 * KiB, 1024),
]
for name, nrows, seg, rs, t0, t1, outer, th in cases:
    for smem in (0, nrows * seg + 4096):  # free occupancy / the occupancy of a transform kernel holding the tile in shared memory
        rc = g(V.data_ptr(), Wv.data_ptr(), nrows, seg, rs, t0, t1, outer, th, smem, s)
        if rc:
            print(name, "not runnable", rc)
            continue
        ms = timeit(lambda: g(V.data_ptr(), Wv.data_ptr(), nrows, seg, rs, t0, t1, outer, th, smem, s), 5)
        print(f"seg copy {name} smem {smem // 1024:3d} KiB: {ms:.4f} ms  {vb / ms / 1e6:7.0f} GB/s  {vb / ms / 1e6 / 6527.8 * 100:5.1f}% of the HBM copy peak")
