"""Measurement aid: what the host can do for pageable arrays -- multi-threaded memcpy into pinned memory, first-touch cost,
cudaHostRegister rate -- next to the PCIe link (tools/sessions).  Usage: python tools/host_copy_probe.py"""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

GB = 1 << 30
n = 2 * GB
src = np.empty(n, dtype=np.uint8)
src[:] = 1
pin = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
pin[:] = 0
dev = torch.empty(n, dtype=torch.uint8, device="cuda")


def par_copy(dst, s, threads, chunk=32 << 20):
    parts = [(o, min(o + chunk // threads, len(s))) for o in range(0, len(s), chunk // threads)]
    with ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter()
        list(ex.map(lambda p: np.copyto(dst[p[0]:p[1]], s[p[0]:p[1]]), parts))
        return time.perf_counter() - t0


for th in (1, 2, 4, 8, 12, 16):
    par_copy(pin, src, th)
    dt = par_copy(pin, src, th)
    print(f"memcpy pageable -> pinned, {th:2d} threads: {n / dt / 1e9:6.1f} GB/s", flush=True)
for th in (4, 8, 16):
    dt = par_copy(src, pin, th)
    print(f"memcpy pinned -> pageable, {th:2d} threads: {n / dt / 1e9:6.1f} GB/s", flush=True)
t = torch.from_numpy(pin)
torch.cuda.synchronize()
t0 = time.perf_counter()
dev.copy_(t, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D pinned: {n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
fresh = np.empty(n, dtype=np.uint8)
t0 = time.perf_counter()
fresh[:] = 0
print(f"first touch of a fresh 2 GiB array: {(time.perf_counter() - t0) * 1e3:.0f} ms")
rt = torch.cuda.cudart()
for rep in range(2):
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(src.ctypes.data, n, 0)
    t1 = time.perf_counter()
    ts = torch.from_numpy(src)
    dev.copy_(ts, non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rt.cudaHostUnregister(src.ctypes.data)
    t3 = time.perf_counter()
    print(f"cudaHostRegister 2 GiB: rc={rc} {(t1 - t0) * 1e3:.0f} ms, H2D from it {n / (t2 - t1) / 1e9:.1f} GB/s, unregister {(t3 - t2) * 1e3:.0f} ms")
