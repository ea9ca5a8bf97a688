"""Round-2 measurement aids (csrc/probe.cu): bandwidth of an L2-resident working set against a DRAM-sized one, and of
distributed shared memory inside a thread-block cluster.  Usage: python tools/probe_l2_dsmem.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rocket_fft_b200 as R

lib = C.CDLL(R.LIB_PATH)
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
sink = torch.zeros(4, device=dev)
sweep = lib.rfb200_debug_l2_sweep
sweep.restype = C.c_int
sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]


def timeit(fn, reps=5, warm=2):
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


for mb in (8, 16, 32, 64, 96, 256, 2048):
    n = mb << 20
    x = torch.randn(n // 4, device=dev)
    y = torch.empty_like(x)
    reps = max(1, 4096 // mb)
    for mode, name in ((0, "read"), (1, "copy")):
        for ctas in (296, 592):
            ms = timeit(lambda: sweep(x.data_ptr(), y.data_ptr(), n, reps, mode, ctas, sink.data_ptr(), s))
            moved = n * reps * (2 if mode else 1)
            print(f"L2 sweep {name} {mb:5d} MiB x{reps:4d} reps, {ctas} CTAs: {ms:8.3f} ms  {moved / ms / 1e6:8.0f} GB/s", flush=True)
    del x, y

ds = lib.rfb200_debug_dsmem
ds.restype = C.c_int
ds.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
for cluster in (2, 4, 8):
    kb = min(16, 200 // cluster)
    ncl = 148 // cluster
    cyc = torch.zeros(cluster * ncl, dtype=torch.int64, device=dev)
    for mode, name in ((0, "write"), (1, "read")):
        reps = 64
        rc = ds(cluster, ncl, kb, reps, mode, cyc.data_ptr(), sink.data_ptr(), s)
        torch.cuda.synchronize()
        if rc:
            print("dsmem", cluster, name, "rc", rc)
            continue
        ms = timeit(lambda: ds(cluster, ncl, kb, reps, mode, cyc.data_ptr(), sink.data_ptr(), s))
        c = cyc.cpu().double()
        per_cta = kb * 1024 * (cluster - 1) * reps
        print(f"DSMEM {name} cluster {cluster}: {kb} KiB per peer x{reps}: median {c.median().item():9.0f} cyc/CTA -> "
              f"{per_cta / c.median().item():6.1f} B/cyc/SM ({per_cta / c.max().item():6.1f} worst); kernel {ms:.3f} ms -> "
              f"{per_cta * cluster * ncl / ms / 1e6:7.0f} GB/s aggregate over {cluster * ncl} SMs", flush=True)
